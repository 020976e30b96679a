// tests/cxx/msp_yaml.h -- the YAML subset the reference's .msp files are written in, as a small tree parser:
// block maps and block sequences by indentation, `- key: value` items that continue as maps on the following lines, flow
// maps / flow sequences (nested, possibly spanning lines), quoted and plain scalars, `#` comments.  Enough to read the
// reference's own data/config/main-config.msp and data/exemples/*.msp unchanged; no anchors, tags, multi-documents or
// block scalars.  Also: merge of documents the way the reference layers files (maps merge key by key, anything else is
// replaced: src/core/api.cpp:270-281).  Host-only, standard library only (tests/test_frontend_cpu.py runs it without a GPU).
#pragma once
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace msp
{
  struct Node
  {
    enum Kind { Null , Scalar , Map , Seq };
    Kind kind = Null;
    std::string scalar;
    std::vector< std::pair<std::string,Node> > map;      // in file order
    std::vector<Node> seq;

    inline bool is_null() const { return kind == Null; }
    inline bool is_scalar() const { return kind == Scalar; }
    inline bool is_map() const { return kind == Map; }
    inline bool is_seq() const { return kind == Seq; }
    inline const Node* find(const std::string& k) const { if( kind == Map ) for(const auto& kv : map) if( kv.first == k ) return & kv.second; return nullptr; }
    inline Node* find(const std::string& k) { if( kind == Map ) for(auto& kv : map) if( kv.first == k ) return & kv.second; return nullptr; }
    inline Node& set(const std::string& k, const Node& v)
    {
      if( kind != Map ) { *this = Node{}; kind = Map; }
      if( Node* p = find(k) ) { *p = v; return *p; }
      map.push_back( { k , v } ); return map.back().second;
    }
    static inline Node make_scalar(const std::string& s) { Node n; n.kind = Scalar; n.scalar = s; return n; }

    // compact one-line rendering (tests, diagnostics)
    inline std::string str() const
    {
      if( kind == Null ) return "~";
      if( kind == Scalar ) return scalar;
      std::string s = ( kind == Map ) ? "{" : "[";
      bool first = true;
      if( kind == Map ) for(const auto& kv : map) { s += ( first ? "" : "," ) + kv.first + ":" + kv.second.str(); first = false; }
      else for(const auto& v : seq) { s += ( first ? "" : "," ) + v.str(); first = false; }
      return s + ( ( kind == Map ) ? "}" : "]" );
    }
  };

  // b layered on top of a: maps merge key by key (recursively), anything else is replaced
  inline Node merge_nodes(const Node& a, const Node& b)
  {
    if( ! ( a.is_map() && b.is_map() ) ) return b.is_null() ? a : b;
    Node r = a;
    for(const auto& kv : b.map)
    {
      if( Node* p = r.find(kv.first) ) *p = merge_nodes( *p , kv.second );
      else r.map.push_back( kv );
    }
    return r;
  }

  namespace yaml_details
  {
    struct Line { int no; size_t indent; std::string text; };

    inline std::string rtrim(std::string s) { const auto e = s.find_last_not_of(" \t\r\n"); return e == std::string::npos ? std::string() : s.substr(0,e+1); }
    inline std::string strip(std::string s) { const auto b = s.find_first_not_of(" \t\r\n"); return b == std::string::npos ? std::string() : rtrim( s.substr(b) ); }

    // `#` starts a comment when it is outside quotes and at the start of the line or after white space
    inline std::string strip_comment(const std::string& s)
    {
      char q = 0;
      for(size_t i=0;i<s.size();i++)
      {
        const char c = s[i];
        if( q ) { if( c == '\\' && q == '"' ) ++i; else if( c == q ) q = 0; }
        else if( c == '"' || c == '\'' ) q = c;
        else if( c == '#' && ( i == 0 || s[i-1] == ' ' || s[i-1] == '\t' ) ) return s.substr(0,i);
      }
      return s;
    }

    // nesting depth of brackets outside quotes at the end of s (0 = balanced)
    inline int open_brackets(const std::string& s)
    {
      int d = 0; char q = 0;
      for(size_t i=0;i<s.size();i++)
      {
        const char c = s[i];
        if( q ) { if( c == '\\' && q == '"' ) ++i; else if( c == q ) q = 0; }
        else if( c == '"' || c == '\'' ) q = c;
        else if( c == '{' || c == '[' ) ++d;
        else if( c == '}' || c == ']' ) --d;
      }
      return d;
    }

    // position of the `:` that separates key and value of `key: value` (outside quotes and brackets, followed by a blank or
    // the end of the text), or npos
    inline size_t key_colon(const std::string& s)
    {
      int d = 0; char q = 0;
      for(size_t i=0;i<s.size();i++)
      {
        const char c = s[i];
        if( q ) { if( c == '\\' && q == '"' ) ++i; else if( c == q ) q = 0; }
        else if( c == '"' || c == '\'' ) q = c;
        else if( c == '{' || c == '[' ) ++d;
        else if( c == '}' || c == ']' ) --d;
        else if( c == ':' && d == 0 && ( i+1 == s.size() || s[i+1] == ' ' || s[i+1] == '\t' ) ) return i;
      }
      return std::string::npos;
    }

    inline std::string unquote(std::string s)
    {
      s = strip(s);
      if( s.size() >= 2 && ( s.front() == '"' || s.front() == '\'' ) && s.back() == s.front() )
      {
        const bool dq = s.front() == '"';
        s = s.substr(1,s.size()-2);
        if( dq )
        {
          std::string o;
          for(size_t i=0;i<s.size();i++)
          {
            if( s[i] == '\\' && i+1 < s.size() ) { const char e = s[++i]; o += ( e == 'n' ) ? '\n' : ( e == 't' ) ? '\t' : e; }
            else o += s[i];
          }
          s = o;
        }
      }
      return s;
    }

    struct Parser
    {
      std::vector<Line> lines;
      std::string file;
      std::string error;

      inline bool fail(int lineno, const std::string& what)
      {
        if( error.empty() ) { std::ostringstream o; o << file << ":" << lineno << ": " << what; error = o.str(); }
        return false;
      }

      // flow / scalar value held in one string
      inline bool parse_inline(const std::string& text_in, int lineno, Node& out)
      {
        const std::string text = strip(text_in);
        if( text.empty() || text == "~" || text == "null" ) { out = Node{}; return true; }
        if( text.front() == '{' || text.front() == '[' )
        {
          const bool is_map = text.front() == '{';
          if( text.back() != ( is_map ? '}' : ']' ) ) return fail( lineno , "unterminated flow collection" );
          out = Node{}; out.kind = is_map ? Node::Map : Node::Seq;
          const std::string inner = text.substr(1,text.size()-2);
          // split on top-level commas
          std::vector<std::string> parts;
          { int d = 0; char q = 0; std::string cur;
            for(size_t i=0;i<inner.size();i++)
            {
              const char c = inner[i];
              if( q ) { cur += c; if( c == '\\' && q == '"' && i+1 < inner.size() ) cur += inner[++i]; else if( c == q ) q = 0; continue; }
              if( c == '"' || c == '\'' ) { q = c; cur += c; continue; }
              if( c == '{' || c == '[' ) ++d;
              if( c == '}' || c == ']' ) --d;
              if( c == ',' && d == 0 ) { parts.push_back(cur); cur.clear(); continue; }
              cur += c;
            }
            if( ! strip(cur).empty() || ! parts.empty() ) parts.push_back(cur); }
          for(const auto& p : parts)
          {
            if( strip(p).empty() ) return fail( lineno , "empty entry in flow collection" );
            if( is_map )
            {
              const size_t c = key_colon( strip(p) );
              if( c == std::string::npos ) return fail( lineno , "expected `key: value` in flow map" );
              Node v; if( ! parse_inline( strip(p).substr(c+1) , lineno , v ) ) return false;
              out.map.push_back( { unquote( strip(p).substr(0,c) ) , v } );
            }
            else { Node v; if( ! parse_inline( p , lineno , v ) ) return false; out.seq.push_back( v ); }
          }
          return true;
        }
        out = Node::make_scalar( unquote(text) );
        return true;
      }

      // value after `key:` -- inline text, or the block that follows at a deeper indentation (a sequence may sit at the
      // same indentation as its key)
      inline bool parse_value(const std::string& rest, size_t key_indent, size_t& pos, int lineno, Node& out)
      {
        if( ! strip(rest).empty() ) return parse_inline( rest , lineno , out );
        if( pos < lines.size() && ( lines[pos].indent > key_indent || ( lines[pos].indent == key_indent && lines[pos].text[0] == '-' ) ) )
          return parse_block( pos , lines[pos].indent , out );
        out = Node{};
        return true;
      }

      inline bool parse_block(size_t& pos, size_t indent, Node& out)
      {
        out = Node{};
        if( pos >= lines.size() ) return true;
        const bool is_seq = lines[pos].text[0] == '-' && ( lines[pos].text.size() == 1 || lines[pos].text[1] == ' ' );
        out.kind = is_seq ? Node::Seq : Node::Map;
        while( pos < lines.size() && lines[pos].indent == indent )
        {
          const Line ln = lines[pos];
          const bool dash = ln.text[0] == '-' && ( ln.text.size() == 1 || ln.text[1] == ' ' );
          if( dash != is_seq ) { if( is_seq ) break; return fail( ln.no , "sequence item inside a map" ); }
          if( is_seq )
          {
            const std::string item = strip( ln.text.substr(1) );
            const size_t item_indent = indent + ( ln.text.size() - strip(ln.text.substr(1)).size() );    // column of the item's content
            ++pos;
            Node v;
            if( item.empty() )
            {
              if( pos < lines.size() && lines[pos].indent > indent ) { if( ! parse_block( pos , lines[pos].indent , v ) ) return false; }
            }
            else
            {
              const size_t c = ( item.front() == '{' || item.front() == '[' ) ? std::string::npos : key_colon( item );
              if( c == std::string::npos ) { if( ! parse_inline( item , ln.no , v ) ) return false; }
              else
              {
                // `- key: value` opens a map; further keys of the same map sit on the following lines at the content column
                v.kind = Node::Map;
                Node first; if( ! parse_value( item.substr(c+1) , item_indent , pos , ln.no , first ) ) return false;
                v.map.push_back( { unquote( item.substr(0,c) ) , first } );
                if( pos < lines.size() && lines[pos].indent == item_indent && lines[pos].text[0] != '-' )
                {
                  Node more; if( ! parse_block( pos , item_indent , more ) ) return false;
                  for(const auto& kv : more.map) v.map.push_back( kv );
                }
              }
            }
            out.seq.push_back( v );
          }
          else
          {
            const size_t c = key_colon( ln.text );
            if( c == std::string::npos ) return fail( ln.no , "expected `key: value`" );
            ++pos;
            Node v; if( ! parse_value( ln.text.substr(c+1) , indent , pos , ln.no , v ) ) return false;
            out.map.push_back( { unquote( ln.text.substr(0,c) ) , v } );
          }
        }
        if( pos < lines.size() && lines[pos].indent > indent ) return fail( lines[pos].no , "unexpected indentation" );
        return true;
      }
    };
  }

  // parses `text` (named `file` in error messages) into `doc`; on failure returns false and sets `error` to file:line: reason
  inline bool parse_yaml_text(const std::string& text, const std::string& file, Node& doc, std::string& error)
  {
    using namespace yaml_details;
    Parser P; P.file = file;
    std::istringstream in(text);
    std::string raw; int no = 0;
    while( std::getline(in,raw) )
    {
      ++no;
      std::string s = rtrim( strip_comment(raw) );
      if( strip(s).empty() ) continue;
      if( s.find('\t') != std::string::npos && s.find('\t') < s.find_first_not_of(" \t") ) { P.fail( no , "tab in indentation" ); error = P.error; return false; }
      // a flow collection may span several lines: join until the brackets balance
      int startno = no;
      while( open_brackets(s) > 0 && std::getline(in,raw) ) { ++no; s += " " + strip( strip_comment(raw) ); }
      if( open_brackets(s) != 0 ) { P.fail( startno , "unbalanced brackets" ); error = P.error; return false; }
      P.lines.push_back( { startno , s.find_first_not_of(' ') , strip(s) } );
    }
    size_t pos = 0;
    doc = Node{};
    if( P.lines.empty() ) return true;
    if( P.lines[0].indent != 0 ) { P.fail( P.lines[0].no , "document must start at column 0" ); error = P.error; return false; }
    if( ! P.parse_block( pos , 0 , doc ) ) { error = P.error; return false; }
    if( pos != P.lines.size() ) { P.fail( P.lines[pos].no , "unexpected content" ); error = P.error; return false; }
    return true;
  }

  inline bool parse_yaml_file(const std::string& path, Node& doc, std::string& error)
  {
    std::ifstream in(path);
    if( ! in ) { error = "cannot open " + path; return false; }
    std::ostringstream ss; ss << in.rdbuf();
    return parse_yaml_text( ss.str() , path , doc , error );
  }
}
