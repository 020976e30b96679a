// tests/cxx/run_operator.cu -- minimal host for operator plugins: instantiates a registered operator by name, sets slot
// values given as name=value, runs it once (run() = sync, execute(), sync) and dumps any Array2D slot to a file.
// Linked together with UNMODIFIED plugin sources (e.g. the reference's plugins/onikaParallelTutorial/*.cu compiled from
// where they lie) it shows that plugin code runs on this runtime unchanged.
#include <onika/scg/operator.h>
#include <onika/extras/array2d.h>
#include <cstdio>
#include <cstring>
#include <string>

using namespace onika;
using namespace onika::scg;

int main(int argc, char* argv[])
{
  if( argc < 2 )
  {
    std::printf("usage: %s <operator> [slot=value ...] [dump:<array2d slot>=<file>] [fill:<array2d slot>=<rows>x<cols>]\navailable operators:\n", argv[0]);
    for(const auto& n : OperatorNodeFactory::instance()->available_operators()) std::printf("  %s\n", n.c_str());
    return 2;
  }
  auto op = OperatorNodeFactory::instance()->make_operator( argv[1] );
  std::vector< std::pair<std::string,std::string> > dumps;
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
      if( s == nullptr ) { std::fprintf(stderr,"no slot %s\n",key.c_str()+5); return 2; }
      size_t r=0,c=0; std::sscanf( val.c_str() , "%zux%zu" , &r , &c );
      auto* arr = static_cast<extras::Array2D*>( s->value_address() );
      arr->resize(r,c);
      for(size_t k=0;k<r*c;k++)
      { // tests/_inputs.py:lcg_uniform(seed 8, [1e-3,1))
        uint64_t h = ( k * 2654435761ull + 8ull * 40503ull + 12345ull ) & 0xFFFFFFFFull;
        h = ( ( h ^ ( h >> 15 ) ) * 2246822519ull ) & 0xFFFFFFFFull; h = ( h ^ ( h >> 13 ) ) & 0xFFFFFFFFull;
        arr->m_data[k] = 1e-3 + ( 1.0 - 1e-3 ) * ( double(h) / 4294967296.0 );
      }
      continue;
    }
    auto* s = op->slot(key);
    if( s == nullptr || ! s->set_from_string(val) ) { std::fprintf(stderr,"cannot set slot %s\n",key.c_str()); return 2; }
  }
  op->run();
  for(const auto& d : dumps)
  {
    auto* s = op->slot( d.first );
    if( s == nullptr ) { std::fprintf(stderr,"no slot %s\n",d.first.c_str()); return 2; }
    auto* arr = static_cast<extras::Array2D*>( s->value_address() );
    FILE* f = std::fopen( d.second.c_str() , "wb" );
    const uint64_t hdr[2] = { arr->rows() , arr->columns() };
    std::fwrite( hdr , 8 , 2 , f );
    std::fwrite( arr->data() , 8 , arr->rows()*arr->columns() , f );
    std::fclose(f);
  }
  op.reset();
  return 0;
}
