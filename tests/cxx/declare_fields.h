// field declarations for the C++ front-end tests (same declaration pattern plugins use: specialise FieldId<tag>)
#pragma once
#include <cstdint>
#include <onika/soatl/field_id.h>

#define TEST_DECLARE_FIELD(T,NAME,DESC) \
  struct _tag_##NAME {}; \
  namespace onika { namespace soatl { template<> struct FieldId<_tag_##NAME> { using value_type = T; using Id = _tag_##NAME; static const char* name() { return DESC; } }; } } \
  [[maybe_unused]] static constexpr ::onika::soatl::FieldId<_tag_##NAME> NAME {};

TEST_DECLARE_FIELD(double       , particle_rx    , "Particle position X")
TEST_DECLARE_FIELD(double       , particle_ry    , "Particle position Y")
TEST_DECLARE_FIELD(double       , particle_rz    , "Particle position Z")
TEST_DECLARE_FIELD(unsigned char, particle_atype , "Particle atom type")
TEST_DECLARE_FIELD(double       , particle_e     , "Particle energy")
TEST_DECLARE_FIELD(int32_t      , particle_mid   , "Particle molecule id")
TEST_DECLARE_FIELD(float        , particle_dist  , "Particle pair distance")
TEST_DECLARE_FIELD(int16_t      , particle_tmp1  , "Particle Temporary 1")
TEST_DECLARE_FIELD(int8_t       , particle_tmp2  , "Particle Temporary 2")
// config C3: eight fp64 fields swept by one read-modify-write operator
TEST_DECLARE_FIELD(double       , bench_f0       , "C3 field 0")
TEST_DECLARE_FIELD(double       , bench_f1       , "C3 field 1")
TEST_DECLARE_FIELD(double       , bench_f2       , "C3 field 2")
TEST_DECLARE_FIELD(double       , bench_f3       , "C3 field 3")
TEST_DECLARE_FIELD(double       , bench_f4       , "C3 field 4")
TEST_DECLARE_FIELD(double       , bench_f5       , "C3 field 5")
TEST_DECLARE_FIELD(double       , bench_f6       , "C3 field 6")
TEST_DECLARE_FIELD(double       , bench_f7       , "C3 field 7")
