// Prototype of the TMA bulk path considered for the action kernel's gathers / scatter (DESIGN.md
// section 4.1, profiles/r02_action_variants.txt): 1-D cp.async.bulk global -> shared with an
// mbarrier, and cp.reduce.async.bulk shared -> global with .add.f64.  Not used by the engine (the
// runs a warp moves are 25-100 doubles: more issued instructions than the per-lane path); kept so
// that the SASS claim can be checked:
//   nvcc -gencode arch=compute_100a,code=sm_100a -c tools/proto_bulk_copy.cu -o /tmp/p.o
//   cuobjdump -sass /tmp/p.o | grep -E "UBLKCP|UBLKRED|SYNCS"
//     UBLKCP.S.G [UR6], [UR4], UR8 ;   SYNCS.ARRIVE.TRANS64 ... ;   UBLKRED.G.S.ADD.F64.RN [UR4], [UR6], UR7 ;
// Both addresses and the byte count must be multiples of 16: a run of doubles that starts at an odd
// index has to be widened by one element on that side (zero-padded for the reduce).
#include <cstdint>
__device__ __forceinline__ unsigned sa(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const double* x, double* y, int n){
  extern __shared__ __align__(16) unsigned char sm[];
  double* buf = (double*)sm; uint64_t* bar = (uint64_t*)(sm + 4096);
  int lane = threadIdx.x & 31;
  if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sa(bar)), "r"(1)); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  unsigned bytes = 256;
  if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sa(bar)), "r"(bytes) : "memory");
  __syncwarp();
  if (lane == 0) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(sa(buf)), "l"(x), "r"(bytes), "r"(sa(bar)) : "memory");
  unsigned phase = 0;
  asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" :: "r"(sa(bar)), "r"(phase) : "memory");
  buf[lane] *= 2.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" :: "l"(y), "r"(sa(buf)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
