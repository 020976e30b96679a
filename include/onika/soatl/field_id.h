// onika/soatl/field_id.h -- compile-time field identifiers.  A field is declared by specialising FieldId<tag> with
// value_type / Id / name() (see tests/cxx/declare_fields.h); containers are parameterised by lists of tags.
#pragma once
#include <cstdlib>
#include <type_traits>

namespace onika { namespace soatl {

  template<typename _field_id> struct FieldId;
  template<> struct FieldId<void> { using value_type = void; using Id = void; static const char* name() { return "<uknown>"; } };

  static constexpr size_t bad_field_index = 1ull << 30;
  template<typename... ids> struct FieldIds {};

  // position of tag k in a pack (bad_field_index when absent)
  template<typename k, typename... ids> struct find_index_of_id
  {
    static constexpr size_t compute()
    {
      constexpr bool match[] = { std::is_same_v<k,ids>... , false };
      for(size_t i=0; i<sizeof...(ids); i++) if( match[i] ) return i;
      return bad_field_index;
    }
    static constexpr size_t index = compute();
  };
  template<typename k, typename... ids> static inline constexpr size_t find_index_of_id_v = find_index_of_id<k,ids...>::index;
  template<class k, class Fids> struct find_index_of_id_in_field_ids;
  template<class k, class... ids> struct find_index_of_id_in_field_ids< k, FieldIds<ids...> > { static constexpr size_t index = find_index_of_id<k,ids...>::index; };
  template<class k, class Fids> static inline constexpr size_t find_index_of_id_in_field_ids_v = find_index_of_id_in_field_ids<k,Fids>::index;

  template<typename k, typename Fids> struct PrependId;
  template<typename k, typename... ids> struct PrependId<k,FieldIds<ids...> > { using type = FieldIds<k,ids...>; };
  template<typename k, typename Fids> using prepend_field_id_t = typename PrependId<k,Fids>::type;

  // tags that precede k in a list
  template<typename k, typename Fids> struct PrecedingIds;
  template<typename k> struct PrecedingIds<k,FieldIds<> > { using type = FieldIds<>; };
  template<typename k, typename id1, typename... ids> struct PrecedingIds<k,FieldIds<id1,ids...> >
  {
    using type = std::conditional_t< std::is_same_v<k,id1> , FieldIds<> , prepend_field_id_t<id1 , typename PrecedingIds<k,FieldIds<ids...> >::type > >;
  };
  template<typename k, typename... ids> using preceding_field_ids_t = typename PrecedingIds< k, FieldIds<ids...> >::type;

  // first N tags of a list
  template<size_t N, class Fids> struct NFirstFields { using type = FieldIds<>; };
  template<size_t N, class id1, class... ids> struct NFirstFields<N, FieldIds<id1,ids...> >
  {
    using type = std::conditional_t< N==0 , FieldIds<> , prepend_field_id_t<id1 , typename NFirstFields< ( N>0 ? N-1 : 0 ) ,FieldIds<ids...> >::type > >;
  };

  // shortest leading run of FidsRef that contains every tag of a sub-set (a tag absent from FidsRef asks for more than the
  // whole list, i.e. the whole list): FieldIdsIncludingSequence<FidsRef,FieldIds<sub...>>::type / ::n_fields()
  template<class FidsRef , class FidsSubSet> struct FieldIdsIncludingSequence;
  template<class FidsRef , class... SubSetIds>
  struct FieldIdsIncludingSequence< FidsRef , FieldIds<SubSetIds...> >
  {
    static inline constexpr size_t n_fields()
    {
      size_t n = 0;
      for( size_t last : { size_t(0) , ( size_t(1) + find_index_of_id_in_field_ids_v<SubSetIds,FidsRef> ) ... } ) if( last > n ) n = last;
      return n;
    }
    using type = typename NFirstFields< n_fields() , FidsRef >::type;
  };

} }
