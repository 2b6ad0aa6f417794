// force.cuh — the non-bonded force/energy kernels.
//
// brick_force_kernel: the production path. One CTA per brick of cells; the brick's halo (all atoms
// within r_list of any owned atom) is staged into shared memory by the TMA engine (cp.async.bulk +
// mbarrier) as float4 (x,y,z,q) [+ float2 LJ parameters], converted to a brick-local frame, and then
// every owned atom is processed by LPA (=8) lanes that walk its full-shell neighbour list (16-bit halo
// indices, read as coalesced 8-byte words) with four independent pair evaluations in flight per lane.
// Partial forces are reduced with warp shuffles; there are no atomics and no Newton-3 scatter, so
// results are bitwise reproducible run to run. Replaces force_kernel!/energy_kernel!
// (ext/MollyCUDAExt.jl:1595-2045, :2062-2294) and pairwise_force_kernel_nl! (src/kernels.jl:114-140).
//
// allpairs_force_kernel: O(N^2) minimum-image kernel for systems without a usable neighbour list
// (use_neighbors=false / NoCutoff / boxes smaller than 2.5 r_list); replaces
// pairwise_force_kernel_nonl! (ext/MollyCUDAExt.jl:2305-2371).
#pragma once
#include "cells.cuh"
#include "pair.cuh"
#include "peer.cuh"

namespace mb {

constexpr int FORCE_THREADS = 256;
#ifndef MB_LIST_BATCH
#define MB_LIST_BATCH 8
#endif
#ifndef MB_USE_F32X2
#define MB_USE_F32X2 1  // Blackwell packed-f32 (FFMA2/FMUL2) pair loop for the uniform-LJ f32 force path
#endif
#ifndef MB_MIN_BLOCKS
#define MB_MIN_BLOCKS 4
#endif

template <typename T>
struct ForceOut {
    typename VT<T>::T4* f4;   // per-slot force (w unused)
    double* pe_partial;       // [nbricks] (ENERGY)
    double* vir_partial;      // [nbricks*6] xx,yy,zz,xy,xz,yz (ENERGY)
    PeerWait gate;            // decomposed run over peer memory: epoch flags the halo data of this step arrives under
};

// f64 variants get half the resident CTAs (128 registers): under the f32 bound of 64 they spilled 560-1113 LDL/STL each
template <typename T, int COUL, bool UNIFORM, int CUTM, bool ENERGY, int LPA>
__global__ void __launch_bounds__(FORCE_THREADS, (sizeof(T) == 8) ? 2 : MB_MIN_BLOCKS)
    brick_force_kernel(Geom<T> g, PairParams<T> P, const BrickHdr* __restrict__ hdrs, const Run* __restrict__ runs,
                       const IRow* __restrict__ irows, const typename VT<T>::T4* __restrict__ pos4,
                       const typename VT<T>::T2* __restrict__ lj2, const unsigned short* __restrict__ list,
                       const unsigned short* __restrict__ slist, const ushort2* __restrict__ counts, ForceOut<T> out,
                       int brick0) {
    using T4 = typename VT<T>::T4;
    using T2 = typename VT<T>::T2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x + brick0;  // brick0: first brick of this rank's slab (0 on a single GPU)
    const BrickHdr hd = hdrs[b];
    const int tid = threadIdx.x;
    if (hd.i_count == 0 || hd.halo_count > g.halo_cap) {
        if (ENERGY && tid == 0) {
            out.pe_partial[b] = 0.0;
            for (int k = 0; k < 6; k++) out.vir_partial[(size_t)b * 6 + k] = 0.0;
        }
        return;
    }
    if (out.gate.n > 0) {  // the neighbours' drift kernels store this step's halo positions into pos4 (peer.cuh)
        if (tid < out.gate.n) spin_until(out.gate.flag[tid], out.gate.epoch);
        __syncthreads();
        fence_proxy_async_all();  // the TMA engine reads them next
    }
    T4* s_pos = reinterpret_cast<T4*>(smem_raw);
    T2* s_lj = reinterpret_cast<T2*>(s_pos + g.halo_cap);
    uint32_t s_pos_u32 = smem_u32(s_pos);
    asm volatile("" : "+r"(s_pos_u32));  // keep it in a register (otherwise the 5-instruction window-base computation is redone per group)
    __shared__ uint64_t s_bar;
    __shared__ IRow s_rows[64];
    const Run* my_runs = runs + (size_t)b * g.max_runs;
    for (int k = tid; k < g.n_irows; k += blockDim.x) s_rows[k] = irows[(size_t)b * g.n_irows + k];
    stage_halo_issue<T, !UNIFORM>(g, hd, my_runs, pos4, lj2, s_pos, s_lj, &s_bar);  // TMA copies in flight from here

    constexpr int NSUB = FORCE_THREADS / LPA;
    const int sub = tid / LPA, l = tid % LPA;
    T e_acc = (T)0;
    T vir[6] = {(T)0, (T)0, (T)0, (T)0, (T)0, (T)0};
    // Task bookkeeping. A task = one owned atom handled by one group of LPA lanes. The loop below is software-
    // pipelined: the list length and the first LIST_HALF index words of task t+1 are requested while task t is being
    // evaluated, and those of the first task while the halo is still landing, so the global-memory latency of the
    // neighbour-list stream is off the critical path (each CTA only runs ~4 tasks per lane group).
    constexpr int LIST_HALF = (MB_LIST_BATCH >= 2) ? MB_LIST_BATCH / 2 : 1;
    // task -> (slot, shared-memory index): the row search is done once per owned atom into a table instead of once per
    // lane group and iteration (it was 5 % of the kernel's instructions)
    constexpr int MAX_TASKS = 512;
    __shared__ int2 s_task[MAX_TASKS];
    auto search = [&](int task, int& slot, int& si) {
        int q = 0;
        while (q + 1 < g.n_irows && s_rows[q + 1].cum <= task) q++;
        const IRow row = s_rows[q];
        slot = row.slot_begin + (task - row.cum);
        si = row.smem_begin + (task - row.cum);
    };
    const bool tabled = hd.i_count <= MAX_TASKS;
    if (tabled) {
        for (int t = tid; t < hd.i_count; t += blockDim.x) {  // s_rows is visible: stage_halo_issue synchronised the CTA
            int sl, sm;
            search(t, sl, sm);
            s_task[t] = make_int2(sl, sm);
        }
        __syncthreads();
    }
    auto locate = [&](int task, int& slot, int& si) -> bool {
        const bool valid = task < hd.i_count;
        slot = 0;
        si = 0;
        if (valid) {
            if (tabled) {
                const int2 t = s_task[task];
                slot = t.x;
                si = t.y;
            } else {
                search(task, slot, si);
            }
        }
        return valid;
    };
    const int words_in_row = g.stride >> 5;  // groups a row can hold
    const int n_iter = (hd.i_count + NSUB - 1) / NSUB;
    int slot, si;
    bool valid = locate(sub, slot, si);  // s_rows is visible: stage_halo_issue synchronised the CTA
    ushort2 cnt = valid ? counts[slot] : make_ushort2(0, 0);
    uint2 wa[LIST_HALF];
    if (LPA == 8) {
        const uint2* lp2 = reinterpret_cast<const uint2*>(list + (size_t)slot * g.stride) + l;
#pragma unroll
        for (int u = 0; u < LIST_HALF; u++) wa[u] = (valid && u < words_in_row) ? ldg_stream_u2(lp2 + (size_t)u * 8) : make_uint2(0u, 0u);
    }
    stage_halo_wait<T, false>(g, b, hd, my_runs, s_pos, &s_bar);

    for (int it = 0; it < n_iter; it++) {
        const T4 pi = s_pos[si];
        T lj_s_i = (T)0, lj_e_i = (T)0;
        if (!UNIFORM) {
            T2 t = s_lj[si];
            lj_s_i = t.x;
            lj_e_i = t.y;
        }
        const T kq_i = P.ke * pi.w;
        T fx = (T)0, fy = (T)0, fz = (T)0;
#if MB_USE_F32X2
        float2 axx = make_float2(0.f, 0.f), ayy = axx, azz = axx;  // packed-f32 accumulators (two neighbours per lane)
        (void)axx; (void)ayy; (void)azz;
#endif
        auto eval = [&](int j, auto special_tag) {
            constexpr bool SPECIAL = decltype(special_tag)::value;
            const T4 pj = s_pos[j];
            T lj_s_j = (T)0, lj_e_j = (T)0;
            if (!UNIFORM) {
                T2 t = s_lj[j];
                lj_s_j = t.x;
                lj_e_j = t.y;
            }
            const T dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const T r2 = dx * dx + dy * dy + dz * dz;
            T fr, e;
            pair_eval<T, COUL, UNIFORM, CUTM, ENERGY, SPECIAL>(P, r2, lj_s_i, lj_e_i, lj_s_j, lj_e_j, kq_i, pj.w, fr, e);
            const T gx = fr * dx, gy = fr * dy, gz = fr * dz;
            fx += gx;
            fy += gy;
            fz += gz;
            if (ENERGY) {
                e_acc += e;
                vir[0] += dx * gx; vir[1] += dy * gy; vir[2] += dz * gz;
                vir[3] += dx * gy; vir[4] += dx * gz; vir[5] += dy * gz;
            }
        };
        // four neighbours at once, stage by stage, so the four shared-memory loads and the four reciprocal
        // chains are independent and in flight together
        auto eval4 = [&](uint2 w) {
#if defined(MB_ABL) && MB_ABL == 1  // ablation: no pair work at all (staging + list streaming + bookkeeping only)
            fx += __uint_as_float((w.x ^ w.y) & 0x3f000000u);
            return;
#endif
            // entries are byte offsets of float4 records (halo index << LIST_SHIFT)
            int j[4] = {(int)(w.x & 0xffffu), (int)(w.x >> 16), (int)(w.y & 0xffffu), (int)(w.y >> 16)};
#if defined(MB_ABL) && MB_ABL == 2  // ablation: arithmetic without the shared-memory gathers
            j[0] = j[1] = j[2] = j[3] = (int)(threadIdx.x & 7) << LIST_SHIFT;
#endif
            T4 pj[4];
            T2 lj[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                pj[u] = lds_pos(s_pos_u32 + (uint32_t)j[u] * (uint32_t)(sizeof(T4) >> LIST_SHIFT), (T)0);
                if (!UNIFORM)
                    lj[u] = *reinterpret_cast<const T2*>(reinterpret_cast<const char*>(s_lj) + (((size_t)j[u] * sizeof(T2)) >> LIST_SHIFT));
            }
#if MB_USE_F32X2
            if constexpr (std::is_same<T, float>::value && UNIFORM && CUTM == CUTM_PLAIN && !ENERGY && COUL == COUL_NONE) {
                // Blackwell packed-f32 path (FADD2 / FMUL2 / FFMA2): two neighbours per instruction
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    const float4 a = pj[2 * h2], b = pj[2 * h2 + 1];
                    // scalar subtractions write straight into register pairs; packing (a.x, b.x) first costs two moves
                    // per component because the gathered float4s arrive as x,y,z,w quads
                    const float2 dx = make_float2(pi.x - a.x, pi.x - b.x);
                    const float2 dy = make_float2(pi.y - a.y, pi.y - b.y);
                    const float2 dz = make_float2(pi.z - a.z, pi.z - b.z);
                    const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    const float2 iv = make_float2(frcp(r2.x), frcp(r2.y));
                    const float2 i3 = __fmul2_rn(__fmul2_rn(iv, iv), iv);
                    const float2 tt = __ffma2_rn(make_float2(P.uni_A, P.uni_A), i3, make_float2(-P.uni_B, -P.uni_B));
                    float2 fr = __fmul2_rn(tt, __fmul2_rn(i3, iv));
                    fr.x = (r2.x <= P.lj_rc2) ? fr.x : 0.f;
                    fr.y = (r2.y <= P.lj_rc2) ? fr.y : 0.f;
                    axx = __ffma2_rn(fr, dx, axx);
                    ayy = __ffma2_rn(fr, dy, ayy);
                    azz = __ffma2_rn(fr, dz, azz);
                }
                return;
            }
#endif
            T dx[4], dy[4], dz[4], r2[4], fr[4], e[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                dx[u] = pi.x - pj[u].x;
                dy[u] = pi.y - pj[u].y;
                dz[u] = pi.z - pj[u].z;
                r2[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                pair_eval<T, COUL, UNIFORM, CUTM, ENERGY, false>(P, r2[u], lj_s_i, lj_e_i, UNIFORM ? (T)0 : lj[u].x,
                                                                  UNIFORM ? (T)0 : lj[u].y, kq_i, pj[u].w, fr[u], e[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const T gx = fr[u] * dx[u], gy = fr[u] * dy[u], gz = fr[u] * dz[u];
                fx += gx;
                fy += gy;
                fz += gz;
                if (ENERGY) {
                    e_acc += e[u];
                    vir[0] += dx[u] * gx; vir[1] += dy[u] * gy; vir[2] += dz[u] * gz;
                    vir[3] += dx[u] * gy; vir[4] += dx[u] * gz; vir[5] += dy[u] * gz;
                }
            }
        };

        // main list: groups of 32 entries; with LPA lanes each lane owns 32/LPA entries per group
        const int n_groups = ((int)cnt.x + 31) >> 5;
        const unsigned short* lp = list + (size_t)slot * g.stride;
        // next task (if any): its list length is requested now, its first index words after the first half below
        int nslot = 0, nsi = 0;
        const bool nvalid = (it + 1 < n_iter) ? locate((it + 1) * NSUB + sub, nslot, nsi) : false;
        ushort2 ncnt = nvalid ? counts[nslot] : make_ushort2(0, 0);
        if (LPA == 8) {
            const uint2* lp2 = reinterpret_cast<const uint2*>(lp) + l;
            // second half of this task's first batch
            uint2 wb[LIST_HALF];
#pragma unroll
            for (int u = 0; u < LIST_HALF; u++)
                wb[u] = (LIST_HALF + u < n_groups) ? ldg_stream_u2(lp2 + (size_t)(LIST_HALF + u) * 8) : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < LIST_HALF; u++)
                if (u < n_groups) eval4(wa[u]);
            // prefetch the next task's first half into the registers just consumed
            {
                const uint2* np2 = reinterpret_cast<const uint2*>(list + (size_t)nslot * g.stride) + l;
#pragma unroll
                for (int u = 0; u < LIST_HALF; u++)
                    wa[u] = (nvalid && u < words_in_row) ? ldg_stream_u2(np2 + (size_t)u * 8) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < LIST_HALF; u++)
                if (LIST_HALF + u < n_groups) eval4(wb[u]);
            // rows longer than one batch
            for (int g0 = 2 * LIST_HALF; g0 < n_groups; g0 += LIST_HALF) {
                uint2 wc[LIST_HALF];
#pragma unroll
                for (int u = 0; u < LIST_HALF; u++)
                    wc[u] = (g0 + u < n_groups) ? ldg_stream_u2(lp2 + (size_t)(g0 + u) * 8) : make_uint2(0u, 0u);
#pragma unroll
                for (int u = 0; u < LIST_HALF; u++)
                    if (g0 + u < n_groups) eval4(wc[u]);
            }
        } else {
            // generic: logical entry m of a group lives at ((m & 7) << 2) + (m >> 3)
            for (int gi = 0; gi < n_groups; gi++) {
                for (int m = l; m < 32; m += LPA) {
                    int phys = ((m & 7) << 2) + (m >> 3);
                    eval((int)lp[gi * 32 + phys] >> LIST_SHIFT, std::false_type{});
                }
            }
        }
        // special (1-4) pairs
        for (int m = l; m < (int)cnt.y; m += LPA) eval((int)slist[(size_t)slot * g.sstride + m], std::true_type{});
#if MB_USE_F32X2
        if constexpr (std::is_same<T, float>::value) {
            fx += axx.x + axx.y;
            fy += ayy.x + ayy.y;
            fz += azz.x + azz.y;
        }
#endif
        // reduce the LPA partial forces
        __syncwarp();
#pragma unroll
        for (int o = LPA >> 1; o > 0; o >>= 1) {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
        }
        if (l == 0 && valid) out.f4[slot] = make4<T>(fx, fy, fz, (T)0);
        slot = nslot; si = nsi; valid = nvalid; cnt = ncnt;
    }
    if (ENERGY) {
        // full shell: every pair was visited from both ends -> 1/2
        __shared__ double s_red[FORCE_THREADS / 32][7];
        double v[7] = {(double)e_acc, (double)vir[0], (double)vir[1], (double)vir[2],
                       (double)vir[3], (double)vir[4], (double)vir[5]};
#pragma unroll
        for (int k = 0; k < 7; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        }
        const int lane = tid & 31, wid = tid >> 5;
        if (lane == 0)
            for (int k = 0; k < 7; k++) s_red[wid][k] = v[k];
        __syncthreads();
        if (tid < 7) {
            double s = 0.0;
            for (int w = 0; w < FORCE_THREADS / 32; w++) s += s_red[w][tid];
            s *= 0.5;
            if (tid == 0) out.pe_partial[b] = s;
            else out.vir_partial[(size_t)b * 6 + (tid - 1)] = s;
        }
    }
}

// deterministic final reduction of per-CTA partials: pe_out[0] += sum, vir_out (3x3, T) += sum
template <typename T>
__global__ void reduce_partials_kernel(int n, const double* __restrict__ pe_partial, const double* __restrict__ vir_partial,
                                       T* pe_out, T* vir_out, double* pe_out_d) {
    __shared__ double s_red[8][7];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += blockDim.x) {
        v[0] += pe_partial[i];
        if (vir_partial)
            for (int k = 0; k < 6; k++) v[1 + k] += vir_partial[(size_t)i * 6 + k];
    }
    for (int k = 0; k < 7; k++)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0)
        for (int k = 0; k < 7; k++) s_red[wid][k] = v[k];
    __syncthreads();
    if (tid == 0) {
        double s[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int w = 0; w < (int)(blockDim.x >> 5); w++)
            for (int k = 0; k < 7; k++) s[k] += s_red[w][k];
        if (pe_out) pe_out[0] += (T)s[0];
        if (pe_out_d) pe_out_d[0] = s[0];
        if (vir_out) {
            // column-major 3x3: W[a,b] += dr[a] f[b]; symmetric here
            vir_out[0] += (T)s[1]; vir_out[4] += (T)s[2]; vir_out[8] += (T)s[3];
            vir_out[1] += (T)s[4]; vir_out[3] += (T)s[4];
            vir_out[2] += (T)s[5]; vir_out[6] += (T)s[5];
            vir_out[5] += (T)s[6]; vir_out[7] += (T)s[6];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// All-pairs path (no neighbour list). Atoms in original order; minimum image by the reference's
// select chain (src/spatial.jl:491-500). One thread per i atom, j tiles through shared memory.
// Exclusions / specials through the per-atom CSR partner lists.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T mic_1d(T ci, T cj, T L) {
    // vector_1D(c1=ci, c2=cj): returns c_j - c_i minimum image
    T v = cj - ci;
    T vp = v + L, vm = v - L;
    return (v > (T)0) ? ((v < -vm) ? v : vm) : ((-v < vp) ? v : vp);
}

constexpr int AP_THREADS = 128;

template <typename T, int COUL, int CUTM, bool ENERGY>
__global__ void __launch_bounds__(AP_THREADS)
    allpairs_force_kernel(int n, PairParams<T> P, T Lx, T Ly, T Lz, const typename VT<T>::T4* __restrict__ posq,
                          const typename VT<T>::T2* __restrict__ lj2, const int* __restrict__ ex_ptr,
                          const int* __restrict__ ex_idx, const int* __restrict__ sp_ptr,
                          const int* __restrict__ sp_idx, typename VT<T>::T4* __restrict__ f4,
                          double* __restrict__ pe_partial, double* __restrict__ vir_partial) {
    using T4 = typename VT<T>::T4;
    using T2 = typename VT<T>::T2;
    __shared__ T4 s_pos[AP_THREADS];
    __shared__ T2 s_lj[AP_THREADS];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * AP_THREADS + tid;
    const bool active = i < n;
    T4 pi = active ? posq[i] : make4<T>(0, 0, 0, 0);
    T2 li = active ? lj2[i] : make2<T>(0, 0);
    const T kq_i = P.ke * pi.w;
    int ex_a = 0, ex_n = 0, sp_a = 0, sp_n = 0;
    if (active && ex_ptr) { ex_a = ex_ptr[i]; ex_n = ex_ptr[i + 1] - ex_a; }
    if (active && sp_ptr) { sp_a = sp_ptr[i]; sp_n = sp_ptr[i + 1] - sp_a; }
    T fx = 0, fy = 0, fz = 0, e_acc = 0;
    T vir[6] = {0, 0, 0, 0, 0, 0};
    for (int j0 = 0; j0 < n; j0 += AP_THREADS) {
        int jj = j0 + tid;
        __syncthreads();
        s_pos[tid] = (jj < n) ? posq[jj] : make4<T>(0, 0, 0, 0);
        s_lj[tid] = (jj < n) ? lj2[jj] : make2<T>(0, 0);
        __syncthreads();
        int lim = min(AP_THREADS, n - j0);
        if (!active) continue;
        for (int k = 0; k < lim; k++) {
            int j = j0 + k;
            if (j == i) continue;
            bool excluded = false, special = false;
            for (int m = 0; m < ex_n; m++) excluded |= (ex_idx[ex_a + m] == j);
            for (int m = 0; m < sp_n; m++) special |= (sp_idx[sp_a + m] == j);
            T4 pj = s_pos[k];
            T2 lj = s_lj[k];
            // d = c_i - c_j = -vector(c_i, c_j)
            T dx = -mic_1d(pi.x, pj.x, Lx), dy = -mic_1d(pi.y, pj.y, Ly), dz = -mic_1d(pi.z, pj.z, Lz);
            T r2 = dx * dx + dy * dy + dz * dz;
            T fr, e;
            pair_eval_rt<T, COUL, CUTM, ENERGY>(P, r2, li.x, li.y, lj.x, lj.y, kq_i, pj.w, excluded, special, fr, e);
            T gx = fr * dx, gy = fr * dy, gz = fr * dz;
            fx += gx; fy += gy; fz += gz;
            if (ENERGY) {
                e_acc += e;
                vir[0] += dx * gx; vir[1] += dy * gy; vir[2] += dz * gz;
                vir[3] += dx * gy; vir[4] += dx * gz; vir[5] += dy * gz;
            }
        }
    }
    if (active) f4[i] = make4<T>(fx, fy, fz, (T)0);
    if (ENERGY) {
        __shared__ double s_red[AP_THREADS / 32][7];
        double v[7] = {(double)e_acc, (double)vir[0], (double)vir[1], (double)vir[2],
                       (double)vir[3], (double)vir[4], (double)vir[5]};
        for (int k = 0; k < 7; k++)
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        const int lane = tid & 31, wid = tid >> 5;
        __syncthreads();
        if (lane == 0)
            for (int k = 0; k < 7; k++) s_red[wid][k] = v[k];
        __syncthreads();
        if (tid < 7) {
            double s = 0.0;
            for (int w = 0; w < AP_THREADS / 32; w++) s += s_red[w][tid];
            s *= 0.5;
            if (tid == 0) pe_partial[blockIdx.x] = s;
            else vir_partial[(size_t)blockIdx.x * 6 + (tid - 1)] = s;
        }
    }
}

}  // namespace mb
