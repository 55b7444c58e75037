// vio_p2p.cuh — small all-reduce (sum) over NVLink peer memory, issued from inside the library.
//
// The distributed reduced solve (vio_bcr.h, BcrDistPlan) has two real exchange steps per trial step - the W-node interface
// system (465 KB at 8 ranks) and the pose update (480 KB) - plus three scalar sums for the LM decisions.  At these sizes an
// NCCL all-reduce is pure latency (2 (W-1) ring steps, 30-60 us at 8 ranks).  Here every rank owns a symmetric mailbox
// (cudaMalloc + CUDA IPC handles, opened once by every peer); ONE kernel per reduction
//   1. pushes the rank's vector into its slot of every peer's mailbox with 16-byte stores over NVLink / NVSwitch,
//   2. publishes an epoch flag on every peer (st.release.sys after a system-scope fence; the last CTA to finish does it),
//   3. waits for the W-1 flags of its own mailbox (ld.acquire.sys) and sums the W contributions in RANK ORDER -
//      every rank adds the same numbers in the same order, so the result is bitwise identical everywhere (the LM decisions
//      of the ranks must not diverge) and reproducible from run to run.
// Slots are double-buffered by epoch parity: a rank can be at most one reduction ahead of a peer (it needs the peer's flag of
// reduction k to finish k), so the slot it overwrites for k + 2 has been consumed.  Larger vectors (the legacy all-reduce
// of S) stay on NCCL.
#pragma once
#include <cstdint>

#define VIO_P2P_MAX_WORLD 16
#define VIO_P2P_CAP_DOUBLES (96 * 1024)  // per slot
#define VIO_P2P_CTAS 16
#define VIO_P2P_THREADS 512

struct P2pView {
    int rank, world;
    double *slots[VIO_P2P_MAX_WORLD];     // base of rank r's mailbox: [2 parities][world][CAP] doubles
    unsigned *flags[VIO_P2P_MAX_WORLD];   // base of rank r's flag words: [world] (stride 32 words)
    unsigned *err;                        // pinned host word (device-visible): set when a peer's flag does not arrive in time
    unsigned *counter;                    // local: [0] CTAs that finished their pushes, [1] CTAs that finished their sums,
                                          // [2] reductions completed so far (the epoch lives on the device so that a CUDA graph
                                          // can replay the kernel with frozen parameters)
};

__device__ __forceinline__ unsigned p2p_ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void p2p_st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// buf[0..count) := sum over the ranks of their buf, in place.  count <= VIO_P2P_CAP_DOUBLES, buf 16-byte aligned.
__global__ void __launch_bounds__(VIO_P2P_THREADS) k_p2p_allreduce(P2pView pv, double *__restrict__ buf, int count) {
    const int me = pv.rank, W = pv.world;
    // every rank runs the same sequence of reductions: the local count of completed ones + 1 is the same number everywhere.
    // It is bumped by the LAST CTA to leave the kernel, i.e. after every CTA has read it here.
    const unsigned epoch = *reinterpret_cast<volatile unsigned *>(pv.counter + 2) + 1u;
    const int par = epoch & 1;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // double2 part (16-byte aligned buffers) + scalar rest; a thread owns the SAME elements in the push and in the sum
    const int n2 = ((reinterpret_cast<uintptr_t>(buf) & 15) == 0) ? count >> 1 : 0;
    // ---- 1. push
    for (int r = 1; r < W; ++r) {
        const int peer = (me + r) % W;  // start at the neighbour: the ranks do not all hammer rank 0 first
        double *dst = pv.slots[peer] + ((size_t)par * W + me) * VIO_P2P_CAP_DOUBLES;
        for (int i = tid; i < n2; i += nth) reinterpret_cast<double2 *>(dst)[i] = reinterpret_cast<const double2 *>(buf)[i];
        for (int i = 2 * n2 + tid; i < count; i += nth) dst[i] = buf[i];
    }
    // ---- 2. publish: all CTAs' stores are fenced, the last one raises the flags
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(pv.counter, 1u);
        last = done == gridDim.x - 1;
        if (last) *pv.counter = 0;  // for the next reduction (stream-ordered)
    }
    __syncthreads();
    if (last && threadIdx.x < W && (int)threadIdx.x != me) {
        __threadfence_system();
        p2p_st_release_sys(pv.flags[threadIdx.x] + 32 * me, epoch);
    }
    // ---- 3. wait for the peers, sum in rank order
    if (threadIdx.x < W && (int)threadIdx.x != me) {
        const unsigned *f = pv.flags[me] + 32 * threadIdx.x;
        // a rank that died or left the collective sequence must not hang the others' GPUs: give up after ~10 s, tell the host
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(p2p_ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(20);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ull) { *reinterpret_cast<volatile unsigned *>(pv.err) = 1u + threadIdx.x; break; }
        }
    }
    __syncthreads();
    const double *mine = pv.slots[me] + (size_t)par * W * VIO_P2P_CAP_DOUBLES;
    for (int i = tid; i < n2; i += nth) {
        double2 s = make_double2(0.0, 0.0);
        for (int r = 0; r < W; ++r) {
            const double2 x = r == me ? reinterpret_cast<const double2 *>(buf)[i]
                                      : __ldcv(reinterpret_cast<const double2 *>(mine + (size_t)r * VIO_P2P_CAP_DOUBLES) + i);
            s.x += x.x; s.y += x.y;
        }
        reinterpret_cast<double2 *>(buf)[i] = s;
    }
    for (int i = 2 * n2 + tid; i < count; i += nth) {
        double s = 0.0;
        for (int r = 0; r < W; ++r) s += r == me ? buf[i] : __ldcv(mine + (size_t)r * VIO_P2P_CAP_DOUBLES + i);
        buf[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(pv.counter + 1, 1u) == gridDim.x - 1) {
            pv.counter[1] = 0;
            __threadfence();
            pv.counter[2] = epoch;
        }
    }
}
