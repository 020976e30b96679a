// onika/scg/operator_factory.h -- name -> creator registry filled by plugins at load time (ONIKA_AUTORUN_INIT).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <onika/scg/operator.h>

namespace onika { namespace scg {

  struct OperatorCreationException : public std::runtime_error { using std::runtime_error::runtime_error; };

  class OperatorNodeFactory
  {
  public:
    using OperatorCreator = std::function< std::shared_ptr<OperatorNode>() >;
    static inline OperatorNodeFactory* instance() { static OperatorNodeFactory* f = new OperatorNodeFactory(); return f; }
    inline size_t register_factory(const std::string& name, OperatorCreator creator) { m_creators[name] = creator; return m_creators.size(); }
    inline std::shared_ptr<OperatorNode> make_operator(const std::string& name) const
    {
      auto it = m_creators.find(name);
      if( it == m_creators.end() ) throw OperatorCreationException( "Could not find a factory for operator '" + name + "'" );
      auto op = it->second();
      op->set_name(name);
      return op;
    }
    inline std::vector<std::string> available_operators() const { std::vector<std::string> v; for(const auto& kv : m_creators) v.push_back(kv.first); return v; }
  private:
    std::map<std::string,OperatorCreator> m_creators;
  };

  template<class OperatorT> inline std::shared_ptr<OperatorNode> make_simple_operator() { return std::make_shared<OperatorT>(); }
  template<class OperatorT> inline std::shared_ptr<OperatorNode> make_compatible_operator() { return std::make_shared<OperatorT>(); }

} }
