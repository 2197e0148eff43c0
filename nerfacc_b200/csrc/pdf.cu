// pdf.cu -- importance (inverse-transform) sampling along rays and per-ray searchsorted.
//
// replaces (paths relative to /root/reference):
//   nerfacc/cuda/csrc/pdf.cu:97-166   importance_sampling_kernel   (one thread per output sample)
//   nerfacc/cuda/csrc/pdf.cu:168-243  compute_intervels_kernel     (second launch over the samples)
//   nerfacc/cuda/csrc/pdf.cu:247-287  searchsorted_kernel
//   nerfacc/cuda/csrc/pdf.cu:293-456  host wrappers (both importance_sampling overloads, searchsorted)
//
// A ray is resampled by one warp (short rays: the proposal-network shapes) or one CTA (long rays): its
// CDF and edge positions are staged in shared memory once, every thread inverts the CDF for its samples
// with the bound search running out of shared memory, the sample centres stay in shared memory, and the
// same threads then write the edges between neighbouring centres -- one launch and one pass over the
// inputs instead of two launches whose threads each binary-search global memory.  Optionally the same launch also maps the edges from the normalised
// axis to ray distance (t_starts / t_ends of the proposal estimator), which otherwise costs six
// elementwise ATen launches per proposal level.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nerfacc_b200.h"
#include "pdf.cuh"

namespace nfa {

constexpr int kIsThreads = 128;
constexpr int kSearchThreads = 256;
constexpr int64_t kIsSmemFloats = 50 * 1024;  // 200 KB of the 227 KB a CTA may opt in to
constexpr int64_t kWarpRayFloats = 1024;       // rays whose tables fit 4 KB take the warp-per-ray kernel

struct IsParams {
    int32_t n_rays;
    const float* vals;
    const float* cdfs;
    const int64_t* in_packed;   // null: batched input, in_edges per ray
    int64_t in_edges;
    const int64_t* out_packed;  // null: batched output, n_out samples per ray; else samples' (start, count)
    const int64_t* iv_packed;   // intervals' (start, count) when out_packed is given
    int64_t n_out;
    int32_t stratified;
    uint64_t seed, offset;
    float* sample_vals;
    int64_t* sample_ray;
    float* iv_vals;
    int64_t* iv_ray;
    uint8_t* iv_left;
    uint8_t* iv_right;
    // optional s -> t mapping of the edges (batched output only)
    float* t_starts;
    float* t_ends;
    float s_min, s_max;
    int32_t lindisp;
};

// One ray, resampled by `kLanes` cooperating threads (a whole CTA or one warp) whose rank is `tid`.
// kStage: `stage` holds room for this ray's CDF, edge positions and sample centres.  Per-ray counts and
// positions are 32-bit (a ray with 2^31 edges does not exist); only the array offsets are 64-bit.
template <bool kStage, int kLanes, bool kCta>
NFA_D void resample_ray(const IsParams& p, int32_t ray, int tid, float* stage, float bias)
{
    int64_t base;
    int32_t n_in;
    if (p.in_packed) {
        base = p.in_packed[2 * (int64_t)ray];
        n_in = (int32_t)p.in_packed[2 * (int64_t)ray + 1];
    } else {
        base = (int64_t)ray * p.in_edges;
        n_in = (int32_t)p.in_edges;
    }
    int64_t s_base, e_base;
    int32_t n;
    if (p.out_packed) {
        s_base = p.out_packed[2 * (int64_t)ray];
        n = (int32_t)p.out_packed[2 * (int64_t)ray + 1];
        e_base = p.iv_packed[2 * (int64_t)ray];
    } else {
        n = (int32_t)p.n_out;
        s_base = (int64_t)ray * n;
        e_base = s_base + ray;
    }
    if (n <= 0) return;  // uniform over the cooperating threads

    const float* cdf = p.cdfs + base;
    const float* val = p.vals + base;
    float* const out_s = p.sample_vals + s_base;
    float* ts = out_s;
    if (kStage) {
        float* s_cdf = stage;
        float* s_val = stage + n_in;
        for (int32_t i = tid; i < n_in; i += kLanes) {
            s_cdf[i] = __ldg(cdf + i);
            s_val[i] = __ldg(val + i);
        }
        cdf = s_cdf;
        val = s_val;
        ts = stage + 2 * n_in;
        if (kCta) __syncthreads(); else __syncwarp();
    }

    const float quiet_nan = __int_as_float(0x7fc00000);
    const float t_min = n_in > 0 ? val[0] : quiet_nan;
    const float t_max = n_in > 0 ? val[n_in - 1] : quiet_nan;
    const float u_floor = n_in > 0 ? cdf[0] : quiet_nan;
    const float u_ceil = n_in > 0 ? cdf[n_in - 1] : quiet_nan;
    const float u_step = f_div(f_sub(u_ceil, u_floor), (float)n);

    for (int32_t sid = tid; sid < n; sid += kLanes) {
        const float t = n_in > 0 ? is_invert<int32_t>(cdf, val, 0, n_in - 1, is_u<int32_t>(u_floor, u_step, sid, bias))
                                 : quiet_nan;
        ts[sid] = t;
        if (kStage) out_s[sid] = t;
        if (p.sample_ray) p.sample_ray[s_base + sid] = ray;
    }
    // centres visible to the cooperating threads (shared memory, or this CTA's own global writes)
    if (kCta) __syncthreads(); else __syncwarp();

    // edges 0 .. n-1 by the thread of the same index; the thread holding the last one adds the closing edge n
    // (a separate pass over k = n would cost a whole extra loop trip when n is a multiple of the group size)
    float* const out_e = p.iv_vals + e_base;
    for (int32_t k = tid; k < n; k += kLanes) {
        const float e = is_edge<int32_t>(ts, n, k, t_min, t_max);
        out_e[k] = e;
        float e_close = 0.f;
        const bool closes = k == n - 1;
        if (closes) {
            e_close = is_edge<int32_t>(ts, n, n, t_min, t_max);
            out_e[n] = e_close;
        }
        if (p.out_packed) {
            p.iv_ray[e_base + k] = ray;
            p.iv_left[e_base + k] = 1;
            p.iv_right[e_base + k] = k > 0;
            if (closes) {
                p.iv_ray[e_base + n] = ray;
                p.iv_left[e_base + n] = 0;
                p.iv_right[e_base + n] = 1;
            }
        } else if (p.t_starts) {
            const float t = stot(e, p.s_min, p.s_max, p.lindisp != 0);
            p.t_starts[s_base + k] = t;
            if (k > 0) p.t_ends[s_base + k - 1] = t;
            if (closes) p.t_ends[s_base + k] = stot(e_close, p.s_min, p.s_max, p.lindisp != 0);
        }
    }
    if (kStage) {  // before the next ray overwrites the staging area
        if (kCta) __syncthreads(); else __syncwarp();
    }
}

// CTA per ray: long rays (hundreds of edges / samples).  kStage = false searches global memory (rays whose
// tables exceed shared memory).
template <bool kStage>
__global__ void __launch_bounds__(kIsThreads) importance_sampling_kernel(IsParams p)
{
    extern __shared__ float smem[];
    for (int32_t ray = blockIdx.x; ray < p.n_rays; ray += gridDim.x) {
        const float bias = p.stratified ? philox_uniform(p.seed, (uint64_t)(int64_t)ray, p.offset) : 0.5f;
        resample_ray<kStage, kIsThreads, true>(p, ray, threadIdx.x, smem, bias);
    }
}

// Warp per ray, kIsWarps rays in flight per CTA: the proposal-network shapes (tens of edges and samples per
// ray, 10^5..10^6 rays), where a CTA per ray would leave most of its threads idle between two barriers.
constexpr int kIsWarps = 8;
// Stratified: a warp takes `chunk` consecutive rays per turn, lane l draws the jitter of ray l (one Philox4x32-10
// block is ~100 instructions; drawn per ray by all 32 lanes it was a quarter of the kernel), then the rays are
// resampled one after the other with the jitter handed over by shuffle.
template <bool kStratified>
__global__ void __launch_bounds__(kIsWarps * 32) importance_sampling_warp_kernel(IsParams p, int32_t floats_per_ray,
                                                                                 int32_t chunk /* rays per turn, <= 32 */)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* stage = smem + (size_t)warp * floats_per_ray;
    if constexpr (!kStratified) {  // one ray per turn, neighbouring rays on neighbouring warps
        for (int64_t ray = (int64_t)blockIdx.x * kIsWarps + warp; ray < p.n_rays; ray += (int64_t)gridDim.x * kIsWarps)
            resample_ray<true, 32, false>(p, (int32_t)ray, lane, stage, 0.5f);
    } else
    for (int64_t ray0 = ((int64_t)blockIdx.x * kIsWarps + warp) * chunk; ray0 < p.n_rays;
         ray0 += (int64_t)gridDim.x * kIsWarps * chunk) {
        const int n = (int)min((int64_t)chunk, p.n_rays - ray0);
        float jitter = 0.5f;
        if (lane < n) jitter = philox_uniform(p.seed, (uint64_t)(ray0 + lane), p.offset);
        for (int i = 0; i < n; ++i)
            resample_ray<true, 32, false>(p, (int32_t)(ray0 + i), lane, stage, __shfl_sync(0xffffffffu, jitter, i));
    }
}

// ---------------------------------------------------------------------------
// Batched input AND output (the proposal-network shapes, BASELINE config 4: [n_rays, 65] -> 32 samples): the hot
// flavour.  Same arithmetic as resample_ray (pdf.cuh), but
//   * a warp takes `chunk` consecutive rays and keeps TWO staging buffers: the CDF / edge rows of ray i+1 are copied
//     global -> shared with cp.async (LDGSTS, no register round trip, no scoreboard wait) while ray i is resampled;
//   * every per-ray quantity that the generic kernel derives from the packed / batched switch is a pointer that
//     advances by a constant.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_f32(float* dst_smem, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

template <bool kStratified, bool kMapT>
__global__ void __launch_bounds__(kIsWarps * 32) importance_sampling_batched_kernel(IsParams p, int32_t n_in, int32_t n,
                                                                                   int32_t chunk)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: stage[2][2 * n_in] (cdf row, then edge-position row) and the n sample centres
    float* const base = smem + (size_t)warp * (4 * n_in + n);
    float* const ts = base + 4 * n_in;
    const float quiet_nan = __int_as_float(0x7fc00000);

    auto stage_ray = [&](int64_t ray, int buf) {  // asynchronous: completes at the next wait
        const float* c = p.cdfs + ray * n_in;
        const float* v = p.vals + ray * n_in;
        float* d = base + buf * 2 * n_in;
        for (int i = lane; i < n_in; i += 32) {
            cp_async_f32(d + i, c + i);
            cp_async_f32(d + n_in + i, v + i);
        }
        cp_async_commit();
    };

    for (int64_t ray0 = ((int64_t)blockIdx.x * kIsWarps + warp) * chunk; ray0 < p.n_rays;
         ray0 += (int64_t)gridDim.x * kIsWarps * chunk) {
        const int cnt = (int)min((int64_t)chunk, p.n_rays - ray0);
        float jitter = 0.5f;
        if (kStratified && lane < cnt) jitter = philox_uniform(p.seed, (uint64_t)(ray0 + lane), p.offset);
        stage_ray(ray0, 0);
        float* out_s = p.sample_vals + ray0 * n;
        float* out_e = p.iv_vals + ray0 * (n + 1);
        float* out_a = kMapT ? p.t_starts + ray0 * n : nullptr;
        float* out_b = kMapT ? p.t_ends + ray0 * n : nullptr;
        for (int i = 0; i < cnt; ++i) {
            const int buf = i & 1;
            if (i + 1 < cnt) {
                stage_ray(ray0 + i + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const float* cdf = base + buf * 2 * n_in;
            const float* val = cdf + n_in;
            const float bias = kStratified ? __shfl_sync(0xffffffffu, jitter, i) : 0.5f;
            const float t_min = val[0], t_max = val[n_in - 1];
            const float u_floor = cdf[0], u_ceil = cdf[n_in - 1];
            const float u_step = f_div(f_sub(u_ceil, u_floor), (float)n);
            for (int sid = lane; sid < n; sid += 32) {
                const float t = is_invert<int32_t>(cdf, val, 0, n_in - 1, is_u<int32_t>(u_floor, u_step, sid, bias));
                ts[sid] = t;
                out_s[sid] = t;
                if (p.sample_ray) p.sample_ray[(ray0 + i) * n + sid] = ray0 + i;
            }
            __syncwarp();
            for (int k = lane; k < n; k += 32) {
                const float e = is_edge<int32_t>(ts, n, k, t_min, t_max);
                out_e[k] = e;
                const bool closes = k == n - 1;
                float e_close = 0.f;
                if (closes) {
                    e_close = is_edge<int32_t>(ts, n, n, t_min, t_max);
                    out_e[n] = e_close;
                }
                if (kMapT) {
                    const float t = stot(e, p.s_min, p.s_max, p.lindisp != 0);
                    out_a[k] = t;
                    if (k > 0) out_b[k - 1] = t;
                    if (closes) out_b[k] = stot(e_close, p.s_min, p.s_max, p.lindisp != 0);
                }
            }
            __syncwarp();  // `ts` and this stage are free again
            out_s += n;
            out_e += n + 1;
            if (kMapT) {
                out_a += n;
                out_b += n;
            }
        }
    }
    (void)quiet_nan;
}

struct SearchParams {
    int64_t n_query;
    const float* q_vals;
    const int64_t* q_packed;  // null: batched query, q_edges per ray
    const int64_t* q_ray;     // optional ray id per query item (flattened query)
    int32_t n_rays;
    int64_t q_edges;
    const float* k_vals;
    const int64_t* k_packed;  // null: batched key, k_edges per ray
    int64_t k_edges;
    int64_t* ids_left;
    int64_t* ids_right;
};

__global__ void __launch_bounds__(kSearchThreads) searchsorted_kernel(SearchParams p)
{
    for (int64_t i = (int64_t)blockIdx.x * kSearchThreads + threadIdx.x; i < p.n_query;
         i += (int64_t)gridDim.x * kSearchThreads) {
        int64_t ray;
        if (!p.q_packed) ray = i / p.q_edges;
        else if (p.q_ray) ray = p.q_ray[i];
        else ray = chunk_upper_bound(p.q_packed, p.n_rays, i) - 1;
        int64_t base, last;
        if (p.k_packed) {
            base = p.k_packed[2 * ray];
            last = base + p.k_packed[2 * ray + 1] - 1;
        } else {
            base = ray * p.k_edges;
            last = base + p.k_edges - 1;
        }
        const int64_t pos = upper_bound_f<int64_t>(p.k_vals, base, last, __ldg(p.q_vals + i));
        int64_t l = pos - 1 < last ? pos - 1 : last;
        if (l < base) l = base;
        int64_t r = pos < last ? pos : last;
        if (r < base) r = base;
        const int64_t rel = p.q_packed ? 0 : base;  // batched queries get per-ray positions
        p.ids_left[i] = l - rel;
        p.ids_right[i] = r - rel;
    }
}

}  // namespace nfa

using namespace nfa;

extern "C" {

int32_t nfa_importance_sampling(int32_t n_rays, const float* vals, const float* cdfs, const int64_t* in_packed_info,
                                int64_t in_edges, int64_t max_in_edges, const int64_t* out_packed_info,
                                const int64_t* iv_packed_info, int64_t n_out, int64_t max_out, int32_t stratified,
                                uint64_t seed, uint64_t offset, float* sample_vals, int64_t* sample_ray_indices,
                                float* iv_vals, int64_t* iv_ray_indices, uint8_t* iv_left, uint8_t* iv_right,
                                float* t_starts, float* t_ends, float s_min, float s_max, int32_t lindisp,
                                nfa_stream_t stream)
{
    if (n_rays < 0 || in_edges < 0 || n_out < 0 || max_in_edges < 0 || max_out < 0) return NFA_ERR_ARG;
    if (n_rays == 0) return NFA_OK;
    if (!sample_vals || !iv_vals) return NFA_ERR_ARG;
    if (max_in_edges > 0 && (!vals || !cdfs)) return NFA_ERR_ARG;
    if (out_packed_info && (!iv_packed_info || !iv_ray_indices || !iv_left || !iv_right)) return NFA_ERR_ARG;
    if ((t_starts != nullptr) != (t_ends != nullptr)) return NFA_ERR_ARG;
    if (t_starts && out_packed_info) return NFA_ERR_UNSUPPORTED;
    if (!out_packed_info && n_out == 0) return NFA_OK;

    IsParams p;
    p.n_rays = n_rays;
    p.vals = vals;
    p.cdfs = cdfs;
    p.in_packed = in_packed_info;
    p.in_edges = in_edges;
    p.out_packed = out_packed_info;
    p.iv_packed = iv_packed_info;
    p.n_out = n_out;
    p.stratified = stratified;
    p.seed = seed;
    p.offset = offset;
    p.sample_vals = sample_vals;
    p.sample_ray = sample_ray_indices;
    p.iv_vals = iv_vals;
    p.iv_ray = iv_ray_indices;
    p.iv_left = iv_left;
    p.iv_right = iv_right;
    p.t_starts = t_starts;
    p.t_ends = t_ends;
    p.s_min = s_min;
    p.s_max = s_max;
    p.lindisp = lindisp;

    const int64_t in_cap = in_packed_info ? max_in_edges : in_edges;
    const int64_t out_cap = out_packed_info ? max_out : n_out;
    const int64_t floats = 2 * in_cap + out_cap;
    cudaStream_t s = (cudaStream_t)stream;
    if (!in_packed_info && !out_packed_info && in_edges >= 2 && n_out >= 1 && 4 * in_edges + n_out <= kWarpRayFloats) {
        // batched rows in, batched rows out (proposal-network shapes): double-buffered warp-per-ray kernel.  A warp
        // walks `chunk` consecutive rays (the next ray's rows are in flight while this one is resampled)
        const int64_t fl = 4 * in_edges + n_out;
        const size_t bytes = (size_t)fl * kIsWarps * sizeof(float);
        int64_t chunk = (int64_t)n_rays / (148 * 8 * kIsWarps);
        chunk = chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk);
        const int64_t want = ((int64_t)n_rays + chunk * kIsWarps - 1) / (chunk * kIsWarps);
        const unsigned grid = (unsigned)(want < (1 << 20) ? want : (1 << 20));
#define NFA_IS_BATCHED(S, M) importance_sampling_batched_kernel<S, M><<<grid, kIsWarps * 32, bytes, s>>>(p, (int32_t)in_edges, (int32_t)n_out, (int32_t)chunk)
        if (stratified) { if (t_starts) NFA_IS_BATCHED(true, true); else NFA_IS_BATCHED(true, false); }
        else { if (t_starts) NFA_IS_BATCHED(false, true); else NFA_IS_BATCHED(false, false); }
#undef NFA_IS_BATCHED
        return (int32_t)cudaGetLastError();
    }
    if (floats <= kWarpRayFloats) {
        // short rays: a warp each; persistent CTAs sized to fill the machine
        const size_t bytes = (size_t)floats * kIsWarps * sizeof(float);
        // stratified: 8 rays per turn, so one Philox block serves 8 rays (117 -> 107 us on config 4; 32 rays per
        // turn left a partial last wave as long as a whole CTA).  Plain: one ray per turn, neighbouring rays on
        // neighbouring warps (92 us against 103 us with turns of 8).
        int64_t chunk = stratified ? (int64_t)n_rays / (148 * 64) : 1;
        chunk = chunk < 1 ? 1 : (chunk > 8 ? 8 : chunk);
        const int64_t want = ((int64_t)n_rays + chunk * kIsWarps - 1) / (chunk * kIsWarps);
        const int64_t cap = 148 * 8 * 4;
        const unsigned grid = (unsigned)(want < cap ? want : cap);
        if (stratified)
            importance_sampling_warp_kernel<true><<<grid, kIsWarps * 32, bytes, s>>>(p, (int32_t)floats, (int32_t)chunk);
        else
            importance_sampling_warp_kernel<false><<<grid, kIsWarps * 32, bytes, s>>>(p, (int32_t)floats, (int32_t)chunk);
        return (int32_t)cudaGetLastError();
    }
    const unsigned grid = (unsigned)(n_rays < (1 << 20) ? n_rays : (1 << 20));
    if (floats <= kIsSmemFloats) {
        const size_t bytes = (size_t)floats * sizeof(float);
        if (bytes > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(importance_sampling_kernel<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e != cudaSuccess) return (int32_t)e;
        }
        importance_sampling_kernel<true><<<grid, kIsThreads, bytes, s>>>(p);
    } else {
        importance_sampling_kernel<false><<<grid, kIsThreads, 0, s>>>(p);
    }
    return (int32_t)cudaGetLastError();
}

int32_t nfa_searchsorted(int64_t n_query, const float* query_vals, const int64_t* query_packed_info,
                         const int64_t* query_ray_indices, int32_t n_rays, int64_t query_edges,
                         const float* key_vals, const int64_t* key_packed_info, int64_t key_edges,
                         int64_t* ids_left, int64_t* ids_right, nfa_stream_t stream)
{
    if (n_query < 0 || n_rays < 0 || query_edges < 0 || key_edges < 0) return NFA_ERR_ARG;
    if (n_query == 0) return NFA_OK;
    if (!query_vals || !key_vals || !ids_left || !ids_right) return NFA_ERR_ARG;
    if (!query_packed_info && query_edges == 0) return NFA_ERR_ARG;
    SearchParams p;
    p.n_query = n_query;
    p.q_vals = query_vals;
    p.q_packed = query_packed_info;
    p.q_ray = query_ray_indices;
    p.n_rays = n_rays;
    p.q_edges = query_edges;
    p.k_vals = key_vals;
    p.k_packed = key_packed_info;
    p.k_edges = key_edges;
    p.ids_left = ids_left;
    p.ids_right = ids_right;
    const int64_t blocks = (n_query + kSearchThreads - 1) / kSearchThreads;
    const unsigned grid = (unsigned)(blocks < 148 * 64 ? blocks : 148 * 64);
    searchsorted_kernel<<<grid, kSearchThreads, 0, (cudaStream_t)stream>>>(p);
    return (int32_t)cudaGetLastError();
}

}  // extern "C"
