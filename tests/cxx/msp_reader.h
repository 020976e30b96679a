// tests/cxx/msp_reader.h -- reader of the flat-batch subset of the reference's .msp (YAML) files used by run_operator.cu:
//     <batch>:
//       name: <string>            (optional)
//       task_group: true|false    (optional)
//       body:
//         - <operator>: { <slot>: <value> , ... }
//         - <operator>
// `#` starts a comment.  Host-only, no dependency besides the standard library (tests/test_frontend_cpu.py runs it here).
#pragma once
#include <cstdio>
#include <fstream>
#include <string>
#include <utility>
#include <vector>

namespace msp
{
  using SlotValues = std::vector< std::pair<std::string,std::string> >;
  struct BodyItem { std::string op; SlotValues values; };
  struct BatchSpec { std::string key, name; bool task_group = false; std::vector<BodyItem> body; };

  inline std::string trim(std::string s)
  {
    const char* ws = " \t\r\n";
    const auto b = s.find_first_not_of(ws);
    if( b == std::string::npos ) return {};
    s = s.substr( b , s.find_last_not_of(ws) - b + 1 );
    if( s.size() >= 2 && ( s.front() == '"' || s.front() == '\'' ) && s.back() == s.front() ) s = s.substr(1,s.size()-2);
    return s;
  }

  // "{ a: 1 , b: x }" -> [(a,1),(b,x)]
  inline bool parse_flow_map(const std::string& text, SlotValues& out)
  {
    const auto lb = text.find('{'), rb = text.rfind('}');
    if( lb == std::string::npos || rb == std::string::npos || rb < lb ) return false;
    std::string inner = text.substr(lb+1,rb-lb-1);
    size_t pos = 0;
    while( pos <= inner.size() )
    {
      const auto comma = inner.find(',',pos);
      const std::string kv = inner.substr( pos , comma == std::string::npos ? std::string::npos : comma-pos );
      const auto colon = kv.find(':');
      if( colon != std::string::npos ) out.push_back( { trim(kv.substr(0,colon)) , trim(kv.substr(colon+1)) } );
      else if( ! trim(kv).empty() ) return false;
      if( comma == std::string::npos ) break;
      pos = comma + 1;
    }
    return true;
  }

  inline bool parse_msp(const std::string& path, std::vector<BatchSpec>& batches)
  {
    std::ifstream in(path);
    if( ! in ) { std::fprintf(stderr,"cannot open %s\n",path.c_str()); return false; }
    std::string line;
    bool in_body = false;
    int lineno = 0;
    while( std::getline(in,line) )
    {
      ++lineno;
      const auto hash = line.find('#');
      if( hash != std::string::npos ) line = line.substr(0,hash);
      if( trim(line).empty() ) continue;
      const size_t indent = line.find_first_not_of(' ');
      const std::string t = trim(line);
      if( indent == 0 )
      {
        if( t.back() != ':' ) { std::fprintf(stderr,"%s:%d: only `<batch>:` definitions are supported at top level\n",path.c_str(),lineno); return false; }
        batches.push_back( BatchSpec{} ); batches.back().key = batches.back().name = trim(t.substr(0,t.size()-1));
        in_body = false;
        continue;
      }
      if( batches.empty() ) { std::fprintf(stderr,"%s:%d: value outside of a batch\n",path.c_str(),lineno); return false; }
      auto& b = batches.back();
      if( t[0] == '-' )
      {
        if( ! in_body ) { std::fprintf(stderr,"%s:%d: list item outside of body\n",path.c_str(),lineno); return false; }
        const std::string item = trim(t.substr(1));
        BodyItem bi;
        const auto colon = item.find(':');
        bi.op = trim( item.substr(0,colon) );
        if( colon != std::string::npos && ! trim(item.substr(colon+1)).empty() && ! parse_flow_map( item.substr(colon+1) , bi.values ) )
        { std::fprintf(stderr,"%s:%d: expected `- op: { slot: value , ... }`\n",path.c_str(),lineno); return false; }
        b.body.push_back( bi );
        continue;
      }
      const auto colon = t.find(':');
      if( colon == std::string::npos ) { std::fprintf(stderr,"%s:%d: expected key: value\n",path.c_str(),lineno); return false; }
      const std::string key = trim(t.substr(0,colon)), val = trim(t.substr(colon+1));
      in_body = false;
      if( key == "name" ) b.name = val;
      else if( key == "task_group" ) b.task_group = ( val == "true" );
      else if( key == "body" ) in_body = true;
      else { std::fprintf(stderr,"%s:%d: unsupported batch attribute '%s'\n",path.c_str(),lineno,key.c_str()); return false; }
    }
    return true;
  }

}
