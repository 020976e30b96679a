// tests/cxx/run_operator.cu -- minimal host for operator plugins.
//   run_operator <operator> [slot=value ...] [fill:<array2d slot>=<rows>x<cols>] [dump:<array2d slot>=<file>]
//       instantiates a registered operator by name, sets slot values, runs it once (run() = sync, execute(), sync)
//   run_operator --msp <file.msp> [batch name] [dump:<slot>=<file> ...]
//       reads the flat-batch subset of the reference's .msp (YAML) files, e.g. data/exemples/concurrent_block_parallel.msp:
//           <batch>:
//             name: <string>            (optional)
//             task_group: true|false    (optional: operators of the body share one parallel execution queue)
//             body:
//               - <operator>: { <slot>: <value> , ... }
//               - <operator>
//       operators run in body order; a consuming slot without a value in the file aliases the latest earlier producing
//       slot of the same name (INPUT_OUTPUT / OUTPUT), which is how `my_array` travels from one operator to the next
// Linked together with UNMODIFIED plugin sources (e.g. the reference's plugins/onikaParallelTutorial/*.cu compiled from
// where they lie) it shows that plugin code runs on this runtime unchanged.
#include <onika/scg/operator.h>
#include <onika/extras/array2d.h>
#include "msp_reader.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

using namespace onika;
using namespace onika::scg;

namespace
{
  using namespace msp;

  // a flat batch: runs its operators in order; with task_group they all enqueue into this node's queue
  class FlatBatch : public OperatorNode
  {
  public:
    std::vector< std::shared_ptr<OperatorNode> > m_ops;
    void execute() override final { for(auto& op : m_ops) op->run(); }
  };

  void fill_array(extras::Array2D* arr, size_t r, size_t c)
  {
    arr->resize(r,c);
    for(size_t k=0;k<r*c;k++)
    { // tests/_inputs.py:lcg_uniform(seed 8, [1e-3,1))
      uint64_t h = ( k * 2654435761ull + 8ull * 40503ull + 12345ull ) & 0xFFFFFFFFull;
      h = ( ( h ^ ( h >> 15 ) ) * 2246822519ull ) & 0xFFFFFFFFull; h = ( h ^ ( h >> 13 ) ) & 0xFFFFFFFFull;
      arr->m_data[k] = 1e-3 + ( 1.0 - 1e-3 ) * ( double(h) / 4294967296.0 );
    }
  }

  bool dump_array(OperatorSlotBase* s, const std::string& file)
  {
    if( s == nullptr || s->value_type() != typeid(extras::Array2D) ) return false;
    auto* arr = static_cast<extras::Array2D*>( s->value_address() );
    FILE* f = std::fopen( file.c_str() , "wb" );
    if( f == nullptr ) return false;
    const uint64_t hdr[2] = { arr->rows() , arr->columns() };
    std::fwrite( hdr , 8 , 2 , f );
    std::fwrite( arr->data() , 8 , arr->rows()*arr->columns() , f );
    std::fclose(f);
    return true;
  }

  int run_msp(int argc, char* argv[])
  {
    std::vector<BatchSpec> specs;
    if( argc < 3 || ! parse_msp( argv[2] , specs ) ) return 2;
    std::string want; SlotValues dumps;
    for(int i=3;i<argc;i++)
    {
      std::string a = argv[i];
      if( a.rfind("dump:",0) == 0 && a.find('=') != std::string::npos ) dumps.push_back( { a.substr(5,a.find('=')-5) , a.substr(a.find('=')+1) } );
      else want = a;
    }
    const BatchSpec* spec = nullptr;
    for(const auto& s : specs) if( ! s.body.empty() && ( want.empty() || s.key == want || s.name == want ) ) { spec = &s; break; }
    if( spec == nullptr ) { std::fprintf(stderr,"no runnable batch%s%s in %s\n", want.empty() ? "" : " named ", want.c_str(), argv[2]); return 2; }

    auto batch = std::make_shared<FlatBatch>();
    batch->set_name( spec->name );
    batch->set_task_group( spec->task_group );
    // slot connection by name, as the reference's graph builder resolves it for a flat batch: a slot that consumes
    // (INPUT / INPUT_OUTPUT) and gets no value from the file aliases the latest earlier slot of that name that produces
    // (OUTPUT / INPUT_OUTPUT); INPUT slots never feed each other
    std::map<std::string,OperatorSlotBase*> producer;
    auto produces = [](const OperatorSlotBase* s) { return ! s->is_private() && ( s->direction() == OUTPUT || s->direction() == INPUT_OUTPUT ); };
    auto consumes = [](const OperatorSlotBase* s) { return ! s->is_private() && ( s->direction() == INPUT || s->direction() == INPUT_OUTPUT ); };
    for(const auto& item : spec->body)
    {
      std::shared_ptr<OperatorNode> op;
      try { op = OperatorNodeFactory::instance()->make_operator( item.op ); }
      catch( const OperatorCreationException& e ) { std::fprintf(stderr,"%s\n",e.what()); return 2; }
      op->set_parent( batch.get() );
      for(const auto& [sname,slot] : op->slots())
      {
        bool valued = false;
        for(const auto& kv : item.values) valued = valued || ( kv.first == sname );
        auto it = producer.find(sname);
        if( ! valued && consumes(slot) && it != producer.end() && ! slot->share_value_of( * it->second ) )
        { std::fprintf(stderr,"slot %s of %s: type differs from its producer\n",sname.c_str(),item.op.c_str()); return 2; }
        if( produces(slot) ) producer[sname] = slot;
      }
      for(const auto& [k,v] : item.values)
      {
        auto* s = op->slot(k);
        if( s == nullptr || ! s->set_from_string(v) ) { std::fprintf(stderr,"cannot set slot %s of %s\n",k.c_str(),item.op.c_str()); return 2; }
      }
      batch->m_ops.push_back( op );
    }
    std::printf("batch %s : %d operator(s)%s\n", spec->name.c_str(), int(batch->m_ops.size()), spec->task_group ? ", task_group" : "");
    batch->run();
    for(const auto& d : dumps)
      if( producer.find(d.first) == producer.end() || ! dump_array( producer[d.first] , d.second ) ) { std::fprintf(stderr,"cannot dump slot %s\n",d.first.c_str()); return 2; }
    batch->m_ops.clear();
    return 0;
  }
}

int main(int argc, char* argv[])
{
  if( argc < 2 )
  {
    std::printf("usage: %s <operator> [slot=value ...] [dump:<array2d slot>=<file>] [fill:<array2d slot>=<rows>x<cols>]\n"
                "       %s --msp <file.msp> [batch] [dump:<array2d slot>=<file>]\navailable operators:\n", argv[0], argv[0]);
    for(const auto& n : OperatorNodeFactory::instance()->available_operators()) std::printf("  %s\n", n.c_str());
    return 2;
  }
  if( std::strcmp( argv[1] , "--msp" ) == 0 ) return run_msp( argc , argv );

  std::shared_ptr<OperatorNode> op;
  try { op = OperatorNodeFactory::instance()->make_operator( argv[1] ); }
  catch( const OperatorCreationException& e ) { std::fprintf(stderr,"%s\n",e.what()); return 2; }
  SlotValues dumps;
  for(int i=2;i<argc;i++)
  {
    std::string a = argv[i];
    const auto eq = a.find('=');
    if( eq == std::string::npos ) { std::fprintf(stderr,"bad argument %s\n",a.c_str()); return 2; }
    std::string key = a.substr(0,eq), val = a.substr(eq+1);
    if( key.rfind("dump:",0) == 0 ) { dumps.push_back( { key.substr(5) , val } ); continue; }
    if( key.rfind("fill:",0) == 0 )
    {
      auto* s = op->slot( key.substr(5) );
      size_t r=0,c=0;
      if( s == nullptr || s->value_type() != typeid(extras::Array2D) || std::sscanf( val.c_str() , "%zux%zu" , &r , &c ) != 2 )
      { std::fprintf(stderr,"cannot fill slot %s\n",key.c_str()+5); return 2; }
      fill_array( static_cast<extras::Array2D*>( s->value_address() ) , r , c );
      continue;
    }
    auto* s = op->slot(key);
    if( s == nullptr || ! s->set_from_string(val) ) { std::fprintf(stderr,"cannot set slot %s\n",key.c_str()); return 2; }
  }
  op->run();
  for(const auto& d : dumps)
    if( ! dump_array( op->slot( d.first ) , d.second ) ) { std::fprintf(stderr,"cannot dump slot %s\n",d.first.c_str()); return 2; }
  op.reset();
  return 0;
}
