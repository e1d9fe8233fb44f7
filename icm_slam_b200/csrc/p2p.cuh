// p2p.cuh -- the exchange steps of a time-segmented sweep (one segment per GPU of a node) done by this library's own kernels
// over peer memory (NVLink / NVSwitch loads and stores into buffers the ranks opened for each other with CUDA IPC), in place
// of the NCCL collectives of multigpu.py:
//
//   far counts     the last block of the association kernel stores (sweep number, far_total) into every rank's window;
//                  k_tail_labels waits for all of them (was: all-gather of 128 B + k_seg_unpack)
//   statistics     k_tail_labels' last block flags this rank's exchange block as final; k_p2p_reduce on rank r waits for every
//                  rank's flag and sums ITS slice of the block (64-bit words [r, r+1) * words / world) over all ranks into its
//                  result buffer -- a reduce-scatter by remote loads; k_fused_means waits for every rank's slice and reads each
//                  landmark's sums from the rank that owns its words -- the all-gather folded into the consumer
//                  (was: all-reduce of the whole block, then the same kernel)
//   halo poses     k_p2p_halo, on the side stream behind the solve: stores this segment's boundary poses into its neighbours'
//                  windows, waits for theirs, fills the halo columns (was: a second all-gather + k_seg_unpack)
//
// Every flag carries the number of the sweep it belongs to (device-resident counters, so a replayed CUDA graph advances them):
// a reader waits for flag >= its own sweep number.  The order of the steps makes a single buffer per flag safe (a rank cannot
// produce sweep k+1's value before every reader of sweep k's has passed), except the halo slots, which are
// double-buffered.  Integer sums: the result does not depend on the order of the ranks, and is bit-identical to the NCCL path.
#pragma once
#include "common.cuh"
#include "assoc.cuh"

#define P2P_MAX_WORLD 16
#define P2P_SPIN_LIMIT (1 << 24)      // polls of a local flag before a rank gives up (ST_P2P_TIMEOUT): seconds, never reached in a healthy job

struct P2PWin {
    unsigned long long far_slot[P2P_MAX_WORLD];   // [source rank] (sweep << 32) | far_total
    unsigned ready[P2P_MAX_WORLD];                // [source rank] sweep whose statistics (incl. new labels) are final there
    unsigned rs_done[P2P_MAX_WORLD];              // [source rank] sweep whose slice has been reduced there
    unsigned long long halo_tag[2][2];            // [parity][0: from the left neighbour, 1: from the right]
    double halo[2][2][8];                         // [parity][side]: left sends (second-to-last, last) poses, right sends its first pose
};

struct P2PDev {
    int on, rank, world;
    P2PWin* const* win;                 // [world] every rank's window (device pointers valid on this GPU)
    const long long* const* exch;       // [world] every rank's exchange block
    const long long* const* res;        // [world] every rank's result buffer (reduced slices)
    long long* my_res;
    long long words, per;               // 64-bit words of the exchange block; words per slice (rank r owns [r * per, (r + 1) * per))
};

__device__ __forceinline__ unsigned long long ld_vol64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
__device__ __forceinline__ unsigned ld_vol32(const unsigned* p) { return *(const volatile unsigned*)p; }

// waits until all `world` 32-bit flags have reached `seq`; false on timeout
__device__ __forceinline__ bool p2p_wait_all32(const unsigned* flags, int world, unsigned seq)
{
    for (int r = 0; r < world; ++r) {
        int it = 0;
        while ((int)(ld_vol32(flags + r) - seq) < 0) { if (++it > P2P_SPIN_LIMIT) return false; }
    }
    return true;
}

// after the far scan (association kernel, last block, first `world` threads): this segment's far_total to every rank
__device__ __forceinline__ void p2p_post_far(const P2PDev& p, unsigned seq, int far_total)
{
    const int r = threadIdx.x;
    if (r < p.world) *(volatile unsigned long long*)&p.win[r]->far_slot[p.rank] = ((unsigned long long)seq << 32) | (unsigned)far_total;
}

// k_tail_labels, one thread per block: wait for every segment's far_total of sweep `seq`; exclusive prefix and total
__device__ __forceinline__ bool p2p_wait_far(const P2PDev& p, unsigned seq, int& base, int& total)
{
    const P2PWin* w = p.win[p.rank];
    base = 0; total = 0;
    for (int r = 0; r < p.world; ++r) {
        unsigned long long v;
        int it = 0;
        while ((int)((unsigned)((v = ld_vol64(&w->far_slot[r])) >> 32) - seq) < 0) { if (++it > P2P_SPIN_LIMIT) return false; }
        const int f = (int)(unsigned)v;
        if (r < p.rank) base += f;
        total += f;
    }
    return true;
}

// this rank's exchange block is final (k_tail_labels, last block, first `world` threads; the caller has fenced)
__device__ __forceinline__ void p2p_post_ready(const P2PDev& p, unsigned seq)
{
    const int r = threadIdx.x;
    if (r < p.world) *(volatile unsigned*)&p.win[r]->ready[p.rank] = seq;
}

// Reduce-scatter by remote loads: my slice of the exchange block summed over all ranks (64-bit integer words: the fixed-point
// sums add exactly; a new label's fp64 mean is non-zero on one rank only, so adding bit patterns reproduces it; two int32
// counts per word never carry).  Then the slice is flagged on every rank.
__global__ void __launch_bounds__(256)
k_p2p_reduce(const P2PDev p, const unsigned* seq_ptr, int* ticket, DevState* st, unsigned long long* trace_row /* or null */)
{
    __shared__ int s_ok;
    if (trace_row && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace_row[((*seq_ptr) & 31) * 8 + 3] = t; }
    const unsigned seq = *(const volatile unsigned*)seq_ptr;
    if (threadIdx.x == 0) {
        s_ok = p2p_wait_all32(p.win[p.rank]->ready, p.world, seq) ? 1 : 0;
        __threadfence_system();      // (acquire: the polling thread fences, the barrier carries it to the block)
    }
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0 && blockIdx.x == 0) st->status |= ST_P2P_TIMEOUT; }
    const long long lo = p.per * p.rank, hi = min(lo + p.per, p.words);
    for (long long w = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; w < hi; w += (long long)gridDim.x * blockDim.x) {
        long long s = 0;
        for (int r = 0; r < p.world; ++r) s += __ldcg(p.exch[r] + w);
        p.my_res[w] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_ok = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_ok) {      // last block: every block's part of the slice is in memory
        if (threadIdx.x == 0) { *ticket = 0; __threadfence_system(); }
        __syncthreads();
        if ((int)threadIdx.x < p.world) *(volatile unsigned*)&p.win[threadIdx.x]->rs_done[p.rank] = seq;
    }
}

// the reduced 64-bit word w of the exchange block, from the rank that owns it
__device__ __forceinline__ long long p2p_word(const P2PDev& p, long long w)
{
    return __ldcg(p.res[(int)(w / p.per)] + w);
}

// Halo poses of the next sweep (side stream, behind the solve; one block of 32 threads).  x = the segment's NEW poses (3 x T, ld).
__global__ void k_p2p_halo(const P2PDev p, unsigned* hseq_ptr, double* __restrict__ x, int64_t ld, int T, int t_lo, int t_hi,
                           double4* __restrict__ ppar, DevState* st, unsigned long long* trace_row /* or null */)
{
    if (trace_row && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace_row[((*hseq_ptr) & 31) * 8 + 6] = t; }
    const unsigned hseq = *(volatile unsigned*)hseq_ptr;
    const int par = hseq & 1, i = threadIdx.x;
    const bool has_l = p.rank > 0, has_r = p.rank + 1 < p.world;
    // push: to the left neighbour my first owned pose (its right halo), to the right one my last two (its left halo)
    if (has_l && i < 3) *(volatile double*)&p.win[p.rank - 1]->halo[par][1][i] = x[i * ld + t_lo];
    if (has_r && i < 3) {
        *(volatile double*)&p.win[p.rank + 1]->halo[par][0][i] = x[i * ld + max(t_hi - 2, t_lo)];
        *(volatile double*)&p.win[p.rank + 1]->halo[par][0][3 + i] = x[i * ld + t_hi - 1];
    }
    __threadfence_system();
    __syncwarp();
    if (i == 0) {
        if (has_l) *(volatile unsigned long long*)&p.win[p.rank - 1]->halo_tag[par][1] = hseq;
        if (has_r) *(volatile unsigned long long*)&p.win[p.rank + 1]->halo_tag[par][0] = hseq;
    }
    // wait for the neighbours' poses of the same sweep, then fill my halo columns (0, 1 and T-1) and their projection parameters
    P2PWin* w = p.win[p.rank];
    bool ok = true;
    if (i == 0) {
        int it = 0;
        if (has_l) while ((long long)(ld_vol64(&w->halo_tag[par][0]) - hseq) < 0) { if (++it > P2P_SPIN_LIMIT) { ok = false; break; } }
        it = 0;
        if (has_r) while ((long long)(ld_vol64(&w->halo_tag[par][1]) - hseq) < 0) { if (++it > P2P_SPIN_LIMIT) { ok = false; break; } }
        if (!ok) st->status |= ST_P2P_TIMEOUT;
    }
    __syncwarp();
    __threadfence_system();
    const volatile double* hl = w->halo[par][0];
    const volatile double* hr = w->halo[par][1];
    if (i < 3) {
        if (has_l) { x[i * ld + 0] = hl[i]; x[i * ld + 1] = hl[3 + i]; }
        if (has_r) x[i * ld + T - 1] = hr[i];
    }
    if (i == 3 && has_l) ppar[0] = make_ppar(hl[0], hl[1], hl[2]);
    if (i == 4 && has_l) ppar[1] = make_ppar(hl[3], hl[4], hl[5]);
    if (i == 5 && has_r) ppar[T - 1] = make_ppar(hr[0], hr[1], hr[2]);
    __syncwarp();
    if (i == 0) *hseq_ptr = hseq + 1u;
}
