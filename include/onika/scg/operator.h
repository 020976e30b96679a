// onika/scg/operator.h -- BOUNDARY SHIM of the operator (plugin) API, just enough for plugins whose execute() drives the
// data-parallel hot path: OperatorNode with parallel_execution_context() / parallel_execution_queue(), typed slots with
// default values (ADD_SLOT), the factory registry and run() = flush+synchronize, execute(), flush+synchronize.
// The reference's graph machinery (full YAML, batch control flow, profiling, MPI) is out of scope;
// this shim keeps plugin SOURCES compiling unchanged and runnable, alone or as a flat batch (tests/cxx/run_operator.cu
// reads the `name: { task_group, body: [ - op: {slot: value} ] }` subset of .msp files, connects same-named slots of the
// same type in body order and lets a task_group batch share one queue among its operators).
#pragma once
#include <cassert>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <onika/log.h>
#include <onika/cpp_utils.h>
#include <onika/cuda/cuda_context.h>
#include <onika/parallel/parallel_execution_context.h>
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/parallel/parallel_execution_operators.h>
#include <onika/scg/operator_slot_direction.h>

namespace onika { namespace scg {

  class OperatorSlotBase;

  class OperatorNode
  {
  public:
    struct DocString { std::string m_doc; };
    struct REQUIRED_t {};  static inline constexpr REQUIRED_t REQUIRED = {};
    struct OPTIONAL_t {};  static inline constexpr OPTIONAL_t OPTIONAL = {};
    static inline constexpr SlotDirection INPUT = ::onika::scg::INPUT;
    static inline constexpr SlotDirection OUTPUT = ::onika::scg::OUTPUT;
    static inline constexpr SlotDirection INPUT_OUTPUT = ::onika::scg::INPUT_OUTPUT;
    struct PRIVATE_t {};   static inline constexpr PRIVATE_t PRIVATE = {};      // ADD_SLOT(T,name,PRIVATE): never connected to other operators

    virtual ~OperatorNode() { m_queue.schedule_all(); m_queue.wait(); }
    virtual void execute() = 0;
    virtual inline bool is_sink() const { return false; }
    virtual inline std::string documentation() const { return {}; }

    inline const std::string& name() const { return m_name; }
    inline void set_name(const std::string& n) { m_name = n; }
    inline void register_slot(const std::string& n, OperatorSlotBase* s) { m_slots[n] = s; }
    inline OperatorSlotBase* slot(const std::string& n) const { auto it = m_slots.find(n); return it == m_slots.end() ? nullptr : it->second; }
    inline const std::map<std::string,OperatorSlotBase*>& slots() const { return m_slots; }

    // ---- glue from the plugin API to the hot path ----
    inline parallel::ParallelExecutionContext* parallel_execution_context(const char* tag = nullptr, const char* sub_tag = nullptr)
    {
      return m_peca.create( tag != nullptr ? tag : m_name.c_str() , sub_tag != nullptr ? sub_tag : "" , & parallel_execution_queue()
                          , m_gpu_enabled ? global_cuda_ctx() : nullptr );
    }
    // operators below a task_group ancestor share that ancestor's queue (src/scg/operator.cpp:563-587)
    inline void set_parent(OperatorNode* p) { m_parent = p; }
    inline OperatorNode* parent() const { return m_parent; }
    inline void set_task_group(bool b) { m_task_group = b; }
    inline bool task_group() const { return m_task_group; }
    inline parallel::ParallelExecutionQueue& parallel_execution_queue()
    {
      for(auto* p = m_parent; p != nullptr; p = p->m_parent) if( p->m_task_group ) return p->parallel_execution_queue();
      if( m_queue.m_stream_pool.m_func == nullptr )
      {
        m_stream_allocator.m_cuda_ctx = global_cuda_ctx();
        m_queue.m_stream_pool = m_stream_allocator.parallel_execution_stream_pool();
      }
      return m_queue;
    }
    inline void set_gpu_enabled(bool b) { m_gpu_enabled = b; }
    static inline onika::cuda::CudaContext* global_cuda_ctx() { return onika::cuda::CudaContext::default_cuda_ctx(); }

    // run = synchronise what previous operators queued, execute, synchronise again
    inline void run()
    {
      parallel_execution_queue() << parallel::flush << parallel::synchronize;
      ONIKA_CU_PROF_RANGE_PUSH( m_name.c_str() );            // one NVTX range per operator (src/scg/operator.cpp:414,487)
      execute();
      parallel_execution_queue() << parallel::flush << parallel::synchronize;
      ONIKA_CU_PROF_RANGE_POP();
    }

  private:
    std::string m_name = "operator";
    std::map<std::string,OperatorSlotBase*> m_slots;
    parallel::ParallelExecutionContextAllocator m_peca;
    parallel::ParallelExecutionStreamAutoAllocator m_stream_allocator;
    parallel::ParallelExecutionQueue m_queue;
    bool m_gpu_enabled = true;
    bool m_task_group = false;
    OperatorNode* m_parent = nullptr;
  };

} }
#include <onika/scg/operator_slot.h>
#include <onika/scg/operator_factory.h>
