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
//   run_operator --app [--config <main-config.msp>] <file.msp> [more.msp ...] [--<key> <value> ...] [dump:<slot>=<file> ...]
//       the reference's application flow (apps/onika-exec.cpp:21-30, src/core/api.cpp:261-312,833) on its own files: the
//       documents are layered (config first, user files on top, `--key value` options last; maps merge key by key), the
//       `configuration` block is set aside, every other top-level key is an operator default / alias / batch definition, and
//       the `simulation` sequence is built into an operator tree: names resolve through aliases (`run: nop`, `--run run_test`),
//       names without a factory become batches (`name`, `task_group`, `body`, nested at will), `- op: {slot: value}` items
//       set slots.  mpi_comm_world / init_cuda / finalize_cuda / global / unit_system / message / nop are stand-ins of the
//       reference's built-in components: the control plane is out of scope, the chain they form is not.
// Linked together with UNMODIFIED plugin sources (e.g. the reference's plugins/onikaParallelTutorial/*.cu compiled from
// where they lie) it shows that plugin code runs on this runtime unchanged.
#include <onika/scg/operator.h>
#include <onika/extras/array2d.h>
#include "msp_reader.h"
#include "msp_yaml.h"
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

  // ---------------------------------------------------------------- application mode (--app)
  // stand-ins of the reference's built-in components named by data/config/main-config.msp (src/scg-builtin-components/*.cpp)
  struct NopNode : public OperatorNode { void execute() override final {} };
  struct MessageNode : public OperatorNode
  {
    ADD_SLOT( std::string , mesg , INPUT , std::string() );
    ADD_SLOT( bool , endl , INPUT , true );
    void execute() override final { std::printf( "%s%s" , mesg->c_str() , *endl ? "\n" : "" ); }
  };
  struct InitCudaNode : public OperatorNode
  {
    void execute() override final
    {
      auto* ctx = global_cuda_ctx();                      // device selection + context: src/scg-builtin-components/init_cuda.cpp:84-225
      if( ctx == nullptr || ! ctx->has_devices() ) fatal_error() << "init_cuda: no CUDA device" << std::endl;
      std::printf( "init_cuda: %s, %d SMs\n" , ctx->m_devices[0].m_deviceProp.name , ctx->m_devices[0].m_deviceProp.multiProcessorCount );
    }
  };
  struct FinalizeCudaNode : public OperatorNode { void execute() override final { ONIKA_B200_CHECK( onk_device_sync() ); std::printf("finalize_cuda\n"); } };

  class BatchNode : public OperatorNode
  {
  public:
    std::vector< std::shared_ptr<OperatorNode> > m_ops;
    void execute() override final { for(auto& op : m_ops) op->run(); }
  };

  struct AppBuilder
  {
    Node defaults;                                        // top-level keys: aliases, operator defaults, batch definitions
    std::map<std::string,OperatorSlotBase*> producer;     // latest producing slot per name, in creation (= execution) order
    std::vector<std::string> trace;                       // flattened tree, for the caller to print / check
    std::string error;

    static bool produces(const OperatorSlotBase* s) { return ! s->is_private() && ( s->direction() == OUTPUT || s->direction() == INPUT_OUTPUT ); }
    static bool consumes(const OperatorSlotBase* s) { return ! s->is_private() && ( s->direction() == INPUT || s->direction() == INPUT_OUTPUT ); }
    bool has_factory(const std::string& n) const { for(const auto& k : OperatorNodeFactory::instance()->available_operators()) if( k == n ) return true; return false; }
    bool fail(const std::string& m) { if( error.empty() ) error = m; return false; }

    bool item(const Node& it, BatchNode* parent, int depth)
    {
      if( it.is_scalar() ) return make( it.scalar , Node{} , parent , depth );
      if( it.is_map() && it.map.size() == 1 ) return make( it.map[0].first , it.map[0].second , parent , depth );
      return fail( "batch body items must be `name` or `name: parameters`, got " + it.str() );
    }

    bool make(std::string name, Node params, BatchNode* parent, int depth)
    {
      if( depth > 32 ) return fail( "operator definitions nest too deep (alias cycle ?) at " + name );
      // aliases: `run: nop`, `--run run_test` (src/scg/operator_factory.cpp:219-240)
      for(int guard=0; guard<32; guard++)
      {
        const Node* d = defaults.find(name);
        if( d != nullptr && d->is_scalar() ) { name = d->scalar; continue; }
        if( params.is_null() && d != nullptr ) params = *d;
        break;
      }
      if( has_factory(name) )
      {
        std::shared_ptr<OperatorNode> op = OperatorNodeFactory::instance()->make_operator( name );
        op->set_parent( parent );
        Node values = params;
        if( params.is_scalar() ) { values = Node{}; values.set( "mesg" , params ); }          // `- message: "text"` (message.cpp yaml_initialize)
        for(const auto& [sname,slot] : op->slots())
        {
          const bool valued = values.is_map() && values.find(sname) != nullptr;
          auto pit = producer.find(sname);
          if( ! valued && consumes(slot) && pit != producer.end() && ! slot->share_value_of( * pit->second ) ) return fail( "slot " + sname + " of " + name + ": type differs from its producer" );
          if( produces(slot) ) producer[sname] = slot;
        }
        if( values.is_map() ) for(const auto& [k,v] : values.map)
        {
          auto* sl = op->slot(k);
          if( sl == nullptr ) { if( op->slots().empty() ) continue; return fail( "operator " + name + " has no slot " + k ); }   // slot-less stand-ins accept any parameters
          if( ! v.is_scalar() || ! sl->set_from_string( v.scalar ) ) return fail( "cannot set slot " + k + " of " + name + " from " + v.str() );
        }
        trace.push_back( std::string( size_t(depth)*2 , ' ' ) + name );
        parent->m_ops.push_back( op );
        return true;
      }
      // no factory: an (implicit) batch, defined by a sequence or by { name, task_group, body }
      if( params.is_null() ) return fail( "Could not find a operator factory for '" + name + "' (and no default declaration exists for this name)" );
      const Node* body = params.is_seq() ? & params : params.find("body");
      if( body == nullptr || ! body->is_seq() ) return fail( "batch '" + name + "' has no body sequence: " + params.str() );
      auto batch = std::make_shared<BatchNode>();
      batch->set_name( name );
      batch->set_parent( parent );
      if( params.is_map() ) for(const auto& [k,v] : params.map)
      {
        if( k == "body" ) continue;
        if( k == "name" && v.is_scalar() ) batch->set_name( v.scalar );
        else if( k == "task_group" && v.is_scalar() ) batch->set_task_group( v.scalar == "true" );
        else return fail( "batch '" + name + "': attribute '" + k + "' is not supported by this runner (control flow is out of scope)" );
      }
      trace.push_back( std::string( size_t(depth)*2 , ' ' ) + name + " [batch " + batch->name() + ( batch->task_group() ? ", task_group]" : "]" ) );
      for(const auto& it : body->seq) if( ! item( it , batch.get() , depth+1 ) ) return false;
      if( parent != nullptr ) parent->m_ops.push_back( batch );
      else root = batch;
      return true;
    }
    std::shared_ptr<BatchNode> root;
  };

  int run_app(int argc, char* argv[])
  {
    auto* F = OperatorNodeFactory::instance();
    for(const char* n : { "nop" , "mpi_comm_world" , "global" , "unit_system" }) F->register_factory( n , make_simple_operator<NopNode> );
    F->register_factory( "message" , make_simple_operator<MessageNode> );
    F->register_factory( "init_cuda" , make_simple_operator<InitCudaNode> );
    F->register_factory( "finalize_cuda" , make_simple_operator<FinalizeCudaNode> );

    Node doc; doc.kind = Node::Map;
    SlotValues dumps;
    std::vector<std::string> files;
    for(int i=2;i<argc;i++)
    {
      std::string a = argv[i];
      if( a == "--config" && i+1 < argc ) { files.insert( files.begin() , argv[++i] ); continue; }
      if( a.rfind("dump:",0) == 0 && a.find('=') != std::string::npos ) { dumps.push_back( { a.substr(5,a.find('=')-5) , a.substr(a.find('=')+1) } ); continue; }
      if( a.rfind("--",0) != 0 ) { files.push_back( a ); continue; }
      // `--a-b value` -> { a: { b: value } } layered on top of the files (src/yaml/cmdline.cpp:54-84); applied after the loop
      if( i+1 < argc && std::string(argv[i+1]).rfind("--",0) != 0 && std::string(argv[i+1]).rfind("dump:",0) != 0 ) ++i;
    }
    for(const auto& f : files)
    {
      Node d; std::string err;
      if( ! parse_yaml_file( f , d , err ) ) { std::fprintf(stderr,"%s\n",err.c_str()); return 2; }
      doc = merge_nodes( doc , d );
    }
    for(int i=2;i<argc;i++)
    {
      std::string a = argv[i];
      if( a == "--config" ) { ++i; continue; }
      if( a.rfind("--",0) != 0 ) continue;
      std::string val = "true";
      if( i+1 < argc && std::string(argv[i+1]).rfind("--",0) != 0 && std::string(argv[i+1]).rfind("dump:",0) != 0 ) val = argv[++i];
      Node v; std::string err;
      if( ! parse_yaml_text( "x: " + val , "<command line>" , v , err ) ) { std::fprintf(stderr,"%s\n",err.c_str()); return 2; }
      Node leaf = v.map[0].second;
      std::string key = a.substr(2);
      for(size_t p = key.rfind('-'); p != std::string::npos; p = key.rfind('-'))
      { Node up; up.set( key.substr(p+1) , leaf ); leaf = up; key = key.substr(0,p); }
      Node top; top.set( key , leaf );
      doc = merge_nodes( doc , top );
    }
    const Node* sim = doc.find("simulation");
    if( sim == nullptr ) { std::fprintf(stderr,"no `simulation` in the loaded files (pass the reference's data/config/main-config.msp with --config)\n"); return 2; }
    AppBuilder B;
    B.defaults.kind = Node::Map;
    for(const auto& kv : doc.map) if( kv.first != "simulation" && kv.first != "configuration" ) B.defaults.map.push_back( kv );
    if( ! B.make( "simulation" , *sim , nullptr , 0 ) ) { std::fprintf(stderr,"%s\n",B.error.c_str()); return 2; }
    std::printf("simulation graph:\n");
    for(const auto& t : B.trace) std::printf("  %s\n", t.c_str());
    B.root->run();
    for(const auto& d : dumps)
      if( B.producer.find(d.first) == B.producer.end() || ! dump_array( B.producer[d.first] , d.second ) ) { std::fprintf(stderr,"cannot dump slot %s\n",d.first.c_str()); return 2; }
    B.root->m_ops.clear();
    B.root.reset();
    return 0;
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
  if( std::strcmp( argv[1] , "--app" ) == 0 ) return run_app( argc , argv );

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
