// onika/scg/operator_slot.h -- typed operator slots: ADD_SLOT(T, name, DIR [, default | REQUIRED | OPTIONAL] [, DocString{..}]).
// Shim: a slot owns its value (default-constructed or initialised from the declared default); values can be set by name
// through OperatorSlotBase (strings for arithmetic types) by a host program.  Slot-to-slot connection is out of scope.
#pragma once
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <onika/scg/operator.h>

namespace onika { namespace scg {

  class OperatorSlotBase
  {
  public:
    virtual ~OperatorSlotBase() = default;
    virtual bool set_from_string(const std::string&) { return false; }
    virtual void* value_address() = 0;
    virtual const std::type_info& value_type() const = 0;
    virtual std::shared_ptr<void> shared_value() const = 0;
    virtual bool share_value_of(const OperatorSlotBase&) { return false; }   // connect: this slot now aliases the other's value
    const std::string& name() const { return m_name; }
    SlotDirection direction() const { return m_dir; }
    bool required() const { return m_required; }
    bool is_private() const { return m_private; }
  protected:
    std::string m_name; SlotDirection m_dir = INPUT_OUTPUT; bool m_required = false; bool m_private = false; std::string m_doc;
  };

  template<class T> using automatic_slot_type_promotion = T;

  template<class T, bool IsInputOnly = false>
  class OperatorSlot : public OperatorSlotBase
  {
  public:
    template<class DirT, class... Args>
    inline OperatorSlot(OperatorNode* op, const char* n, DirT dir, const Args&... args)
    {
      m_name = n; m_value = std::make_shared<T>();
      if constexpr ( std::is_same_v<DirT,OperatorNode::PRIVATE_t> ) { m_dir = INPUT_OUTPUT; m_private = true; }
      else m_dir = dir;
      ( ... , apply(args) );
      if( op != nullptr ) op->register_slot( m_name , this );
    }
    inline T& operator * () const { return *m_value; }
    inline T* operator -> () const { return m_value.get(); }
    inline T* get_pointer() const { return m_value.get(); }
    inline bool has_value() const { return m_value != nullptr; }
    inline void* value_address() override { return m_value.get(); }
    inline const std::type_info& value_type() const override { return typeid(T); }
    inline std::shared_ptr<void> shared_value() const override { return m_value; }
    inline bool share_value_of(const OperatorSlotBase& other) override
    {
      if( other.value_type() != typeid(T) ) return false;
      m_value = std::static_pointer_cast<T>( other.shared_value() );
      return true;
    }
    inline bool set_from_string(const std::string& s) override
    {
      if constexpr ( std::is_same_v<T,bool> )
      {
        if( s == "true" || s == "yes" || s == "on" || s == "1" ) { *m_value = true; return true; }
        if( s == "false" || s == "no" || s == "off" || s == "0" ) { *m_value = false; return true; }
        return false;
      }
      else if constexpr ( std::is_arithmetic_v<T> ) { std::istringstream iss(s); T v{}; if( iss >> v ) { *m_value = v; return true; } return false; }
      else if constexpr ( std::is_same_v<T,std::string> ) { *m_value = s; return true; }
      else return false;
    }
  private:
    inline void apply(const OperatorNode::REQUIRED_t&) { m_required = true; }
    inline void apply(const OperatorNode::OPTIONAL_t&) { m_required = false; }
    inline void apply(const OperatorNode::DocString& d) { m_doc = d.m_doc; }
    template<class U> inline void apply(const U& v) requires std::is_constructible_v<T,const U&> { *m_value = T(v); }
    std::shared_ptr<T> m_value;
  };

  inline constexpr bool operator == (const OperatorNode::PRIVATE_t& , const SlotDirection&) { return false; }

} }

#define ONIKA_SCG_FIRST_ARG_(a,...) a
#define ONIKA_SCG_FIRST_ARG(...) ONIKA_SCG_FIRST_ARG_(__VA_ARGS__,)
#define ADD_SLOT(T,N,...) ::onika::scg::OperatorSlot< ::onika::scg::automatic_slot_type_promotion<T> , ( ONIKA_SCG_FIRST_ARG(__VA_ARGS__) == ::onika::scg::INPUT ) > N { this , #N , __VA_ARGS__ }
