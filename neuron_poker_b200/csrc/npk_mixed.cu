// npk_mixed.cu -- ONE persistent kernel for batches that mix player counts and board sizes (BASELINE config 5: the
// get_equity call of 65,536 self-play tables per step, gym_env/env.py:261-263).
//
// The queries have been sorted by shape on the device (classify kernels in npk_capi.cu): group g = opponents * 6 + known
// board cards holds counts[g] queries, listed in qindex from offsets[g].  Work items of all groups form one sequence --
// group after group -- and the
// warps of a persistent grid (one CTA per SM) pull from it through one global counter.  Per item a warp looks up its
// group and calls that shape's specialised item function (the same code the per-shape kernels inline), so
//   * the rank tables are staged once per CTA per call instead of once per shape,
//   * there is one tail per call instead of one per shape, and no SM partitioning to guess,
//   * the host needs no shape counts: no device->host read, the call is asynchronous and graph-capturable,
//   * neighbouring items belong to the same shape, so the warps of an SM run the same code most of the time.
// Round 1 launched one kernel per shape present (eleven per self-play step, side by side on SM shares sized on the host
// from a synchronous read of the counts): 1.15 ms of Monte-Carlo kernels per step for 65,536 tables.
#include "npk_mc.cuh"

namespace npk {

// Out-of-line so that the persistent kernel stays one dispatch loop plus 60 separately compiled bodies.  The kernel's
// parameter block reaches a body through a pointer (local memory); copying it into locals of the body lets the compiler keep
// the fields the trial loop reads (seed, offsets, counter pointers) in registers -- read through the pointer they were
// reloaded in every iteration (5 LDL + 2 LD per trial in the first version, profiles/r02_ncu_mixed_v1.txt), because the
// shared-memory asm statements of the dealer are memory clobbers.
template <int NOPP, int NB, int REF>
__device__ __noinline__ void shape_item(const EquityParams* p, const WarpCtx* cx, long long q, long long t_begin, long long t_end)
{
    const EquityParams lp = *p;
    const WarpCtx lcx = *cx;
    if (REF) refdeal_item<NOPP, NB>(lp, lcx, q, t_begin, t_end);
    else uniform_item<NOPP, NB>(lp, lcx, q, t_begin, t_end);
}

template <int NOPP, int REF>
__device__ __forceinline__ void dispatch_known(int known, const EquityParams* p, const WarpCtx* cx, long long q, long long b,
                                               long long e)
{
    switch (known) {
        case 0: shape_item<NOPP, 5, REF>(p, cx, q, b, e); break;
        case 1: shape_item<NOPP, 4, REF>(p, cx, q, b, e); break;
        case 2: shape_item<NOPP, 3, REF>(p, cx, q, b, e); break;
        case 3: shape_item<NOPP, 2, REF>(p, cx, q, b, e); break;
        case 4: shape_item<NOPP, 1, REF>(p, cx, q, b, e); break;
        default: shape_item<NOPP, 0, REF>(p, cx, q, b, e); break;
    }
}

template <int REF>
__global__ void __launch_bounds__(kMixedThreads, 1) equity_mixed_kernel(const EquityParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ long long s_first[kShapeGroups + 1];          // first item of every group in the concatenated sequence
    const WarpCtx cx = warp_context(p, smem, 64 + 50 * 32);  // decks sized for the largest shape (preflop: 50 unseen cards)
    const int lane = cx.lane;
    const uint32_t* counts = p.group;                        // [64] queries per group, [64..128) first qindex slot per group
    const long long per_query = p.chunks;
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int g = 0; g < kShapeGroups; g++) { s_first[g] = acc; acc += (long long)counts[g] * per_query; }
        s_first[kShapeGroups] = acc;
    }
    __syncthreads();
    const long long n_items = s_first[kShapeGroups];
    for (long long item = next_item(p.work_counter, lane); item < n_items; item = next_item(p.work_counter, lane)) {
        // group of this item = number of groups that end at or before it (two groups per lane)
        uint32_t before = 0;
        if (lane < kShapeGroups) before += item >= s_first[lane + 1];
        if (lane + 32 < kShapeGroups) before += item >= s_first[lane + 33];
        const int g = (int)__reduce_add_sync(0xffffffffu, before);
        long long qslot, t_begin, t_end;
        item_range(p, item - s_first[g], qslot, t_begin, t_end);
        const long long q = p.qindex[counts[64 + g] + qslot];
        const int known = g % 6;
        switch (g / 6) {
            case 0: dispatch_known<0, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 1: dispatch_known<1, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 2: dispatch_known<2, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 3: dispatch_known<3, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 4: dispatch_known<4, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 5: dispatch_known<5, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 6: dispatch_known<6, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 7: dispatch_known<7, REF>(known, &p, &cx, q, t_begin, t_end); break;
            case 8: dispatch_known<8, REF>(known, &p, &cx, q, t_begin, t_end); break;
            default: dispatch_known<9, REF>(known, &p, &cx, q, t_begin, t_end); break;
        }
    }
}

cudaError_t launch_equity_mixed(const EquityParams& p, int sm_count, cudaStream_t s)
{
    auto k = p.reference_dealer ? equity_mixed_kernel<1> : equity_mixed_kernel<0>;
    const size_t fixed = 128 + (size_t)p.tables.value_bytes + p.tables.rowoff_bytes + p.tables.flush_bytes + kDescBytes;
    const size_t smem = fixed + (size_t)(kMixedThreads / 32) * (64 + 50 * 32) * 4;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<sm_count, kMixedThreads, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace npk
