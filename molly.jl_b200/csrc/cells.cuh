// cells.cuh — device-side neighbour structure: cell binning, deterministic cell sort, brick/halo
// tables and full-shell per-atom neighbour lists. Replaces CellListMap.jl on the host
// (src/neighbors.jl:552-693) and the O((N/32)^2) tile search + O(N^2/1024) mask table of the CUDA
// extension (ext/MollyCUDAExt.jl:1301-1568).
//
// Layout ("slot" = position in the cell-sorted order, x-fastest cell id):
//   pos4[slot]  = (x, y, z, q)            T4   positions are continuous (unwrapped) between rebuilds
//   lj2[slot]   = (sigma part, eps part)  T2   Lorentz: sigma/2, sqrt(eps)  (0,0 if LJ zero shortcut)
//   orig[slot]  = original atom index     int32
// A brick is a box of b[0] x b[1] x b[2] cells owned by one CTA. Its halo is the brick plus h cells on
// every side; every (y,z) row of halo cells is at most 3 contiguous slot runs (periodic wrap in x),
// which is what the force kernel stages into shared memory with 1-D bulk async copies (TMA).
// Neighbour list entries are 16-bit indices into that staged halo.
//
// Every kernel of the rebuild pipeline is gated on ctl->rebuild so the whole sequence can be
// enqueued unconditionally (no host round trip when no atom moved more than skin/2).
#pragma once
#include "common.cuh"

namespace mb {

struct Control {
    int rebuild;        // gate: rebuild requested (set by the drift/ingest kernels or by the host)
    int disp;           // an atom moved more than skin/2 since the last build (fixed-interval policy)
    int overflow;       // bit0 halo capacity, bit1 list stride, bit2 special stride
    int violations;
    int prune;          // gate: the inner (pruned) lists must be refreshed from the outer lists
    int n_prunes;
    unsigned long long n_rebuilds;
    unsigned long long n_pairs;  // real full-shell entries of the last build
    int max_neighbors;
    int max_halo;
    int max_special;
    unsigned int ticket;  // last-block-done counter
    int rebuild_every;    // fixed-interval policy (0 = displacement-triggered)
    long long step;       // MD step counter (simulate!'s step_n), advanced on the device
    long long init_step;  // step_n at the start of the current mb_simulate_vv call
    unsigned int rng[4];  // Andersen thermostat: ctr1 lo/hi, key lo/hi
    // displacement bookkeeping (float bits of squared distances, non-negative floats order like unsigned ints)
    unsigned int max_disp2_bits;       // largest |x - x_ref|^2 since the last rebuild
    unsigned int call_max_disp2_bits;  // largest value any rebuild interval of the current call reached
};

struct BrickHdr {
    int halo_count;  // staged atoms incl. dummy + alignment pads
    int i_count;     // atoms owned by the brick
    unsigned int tx_pos, tx_lj;  // bytes the bulk copies deliver
    int any_shift;
    int pad[3];
};
struct Run {
    int gstart, count, soff, shift;  // shift packed: (wx+1) | (wy+1)<<2 | (wz+1)<<4
};
struct IRow {
    int slot_begin, count, smem_begin, cum;
};

template <typename T>
struct Geom {
    T L[3], invL[3];
    T inv_cell[3];
    double Ld[3], celld[3];
    int nc[3], ncells;
    int b[3], nb[3], nbricks;
    int h, H[3];
    int max_runs, hcells, n_irows;
    int halo_cap, stride, sstride;
    int n;        // atoms
    int align;    // atoms per 16 bytes of the lj2 array (2 for float, 1 for double)
    T rlist2;
    T skin_half2;
    T rinner2;        // dual list: pairs within r_inner at prune time form the list the force kernel walks
    T skin_in_half2;  // ((r_inner - max r_cut) / 2)^2
};

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// ---- R1: wrap into [0,L), cell id, count ------------------------------------------------------
template <typename T>
__global__ void bin_count_kernel(const Control* __restrict__ ctl, Geom<T> g, typename VT<T>::T4* __restrict__ pos4,
                                 int* __restrict__ cid_of, int* __restrict__ cell_count) {
    if (!ctl->rebuild) return;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n) return;
    typename VT<T>::T4 p = pos4[s];
    T x[3] = {p.x, p.y, p.z};
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        T v = x[d] - ffloor(x[d] * g.invL[d]) * g.L[d];  // wrap_coord_1D, src/spatial.jl:573-579
        if (v >= g.L[d]) v -= g.L[d];
        if (v < (T)0) v = (T)0;
        x[d] = v;
        int ci = (int)(v * g.inv_cell[d]);
        c[d] = min(max(ci, 0), g.nc[d] - 1);
    }
    p.x = x[0]; p.y = x[1]; p.z = x[2];
    pos4[s] = p;
    int cid = (c[2] * g.nc[1] + c[1]) * g.nc[0] + c[0];
    cid_of[s] = cid;
    atomicAdd(&cell_count[cid], 1);
}

// ---- R2: exclusive scan of cell counts (single CTA) --------------------------------------------
// Every warp owns a contiguous chunk of cells and walks it in coalesced tiles of 32: chunk sums (lanes accumulate
// independently, one warp reduction) -> scan of the 32 chunk sums -> tile-by-tile shuffle scan with a register carry.
// Two block barriers in total (a tile-by-tile scan across the whole CTA cost four barriers per 1024 cells, 45 us at
// C2's 45 k cells).
__global__ void cell_scan_kernel(const Control* __restrict__ ctl, int ncells, int n, const int* __restrict__ cell_count,
                                 int* __restrict__ cell_start, int* __restrict__ cell_fill) {
    if (!ctl->rebuild) return;
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int per = (((ncells + nw - 1) / nw) + 31) & ~31;  // cells per warp, whole tiles
    const int c0 = min(wid * per, ncells), c1 = min(c0 + per, ncells);
    int sum = 0;
#pragma unroll 4
    for (int c = c0 + lane; c < c1; c += 32) sum += cell_count[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_warp[wid] = sum;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < nw) ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;  // exclusive chunk offsets
    }
    __syncthreads();
    int carry = s_warp[wid];
    for (int base = c0; base < c1; base += 32) {
        const int c = base + lane;
        const int v = (c < c1) ? cell_count[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (c < c1) {
            cell_start[c] = carry + incl - v;
            cell_fill[c] = 0;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) cell_start[ncells] = n;
}

// ---- R3: scatter old slots into their cell segment (order inside the cell fixed up by R4a) -----
__global__ void cell_scatter_kernel(const Control* __restrict__ ctl, int n, const int* __restrict__ cid_of,
                                    const int* __restrict__ cell_start, int* __restrict__ cell_fill,
                                    int* __restrict__ perm) {
    if (!ctl->rebuild) return;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int c = cid_of[s];
    int r = atomicAdd(&cell_fill[c], 1);
    perm[cell_start[c] + r] = s;
}

// ---- R4a: sort each cell segment by old slot (deterministic order), reset counts ---------------
__global__ void cell_sort_kernel(const Control* __restrict__ ctl, int ncells, const int* __restrict__ cell_start,
                                 int* __restrict__ perm, int* __restrict__ cell_count) {
    if (!ctl->rebuild) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int a = cell_start[c], b = cell_start[c + 1];
    for (int i = a + 1; i < b; i++) {
        int v = perm[i];
        int j = i - 1;
        while (j >= a && perm[j] > v) {
            perm[j + 1] = perm[j];
            j--;
        }
        perm[j + 1] = v;
    }
    cell_count[c] = 0;
}

// ---- R4b/R4c: permute the per-slot state into the new order ------------------------------------
template <typename T>
__global__ void permute_gather_kernel(const Control* __restrict__ ctl, int n, const int* __restrict__ perm,
                                      const typename VT<T>::T4* __restrict__ pos4,
                                      const typename VT<T>::T4* __restrict__ vel4,
                                      const typename VT<T>::T2* __restrict__ lj2, const int* __restrict__ orig,
                                      const T* __restrict__ mass, typename VT<T>::T4* __restrict__ pos4_t,
                                      typename VT<T>::T4* __restrict__ vel4_t, typename VT<T>::T2* __restrict__ lj2_t,
                                      int* __restrict__ orig_t, T* __restrict__ mass_t) {
    if (!ctl->rebuild) return;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int o = perm[s];
    pos4_t[s] = pos4[o];
    vel4_t[s] = vel4[o];
    lj2_t[s] = lj2[o];
    orig_t[s] = orig[o];
    mass_t[s] = mass[o];
}
template <typename T>
__global__ void permute_commit_kernel(const Control* __restrict__ ctl, int n,
                                      const typename VT<T>::T4* __restrict__ pos4_t,
                                      const typename VT<T>::T4* __restrict__ vel4_t,
                                      const typename VT<T>::T2* __restrict__ lj2_t, const int* __restrict__ orig_t,
                                      const T* __restrict__ mass_t, typename VT<T>::T4* __restrict__ pos4,
                                      typename VT<T>::T4* __restrict__ vel4, typename VT<T>::T2* __restrict__ lj2,
                                      int* __restrict__ orig, T* __restrict__ mass,
                                      typename VT<T>::T4* __restrict__ xref4, int* __restrict__ inv_orig) {
    if (!ctl->rebuild) return;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    typename VT<T>::T4 p = pos4_t[s];
    pos4[s] = p;
    xref4[s] = p;
    vel4[s] = vel4_t[s];
    lj2[s] = lj2_t[s];
    int o = orig_t[s];
    orig[s] = o;
    inv_orig[o] = s;
    mass[s] = mass_t[s];
}

// ---- R5: brick tables ---------------------------------------------------------------------------
// One CTA per brick. Outputs: hdr, runs[max_runs], irows[n_irows], hcs[hcells] (start,end per halo cell).
template <typename T>
__global__ void brick_tables_kernel(Control* __restrict__ ctl, Geom<T> g, const int* __restrict__ cell_start,
                                    BrickHdr* __restrict__ hdrs, Run* __restrict__ runs, IRow* __restrict__ irows,
                                    ushort2* __restrict__ hcs, int uniform_lj) {
    if (!ctl->rebuild) return;
    extern __shared__ int s_mem[];
    int* s_len = s_mem;                  // max_runs
    int* s_base = s_mem + g.max_runs;    // max_runs
    __shared__ int s_total;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    int B[3] = {b % g.nb[0], (b / g.nb[0]) % g.nb[1], b / (g.nb[0] * g.nb[1])};
    int c0[3], e[3], He[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        c0[d] = B[d] * g.b[d];
        e[d] = min(g.b[d], g.nc[d] - c0[d]);
        He[d] = e[d] + 2 * g.h;
    }
    Run* my_runs = runs + (size_t)b * g.max_runs;
    const int A = g.align;
    // pass 1: run extents
    for (int r = tid; r < g.max_runs; r += blockDim.x) {
        int seg = r % 3, rr = r / 3;
        int ry = rr % g.H[1], rz = rr / g.H[1];
        Run run = {0, 0, 0, 0};
        int len = 0;
        if (ry < He[1] && rz < He[2]) {
            int gy = c0[1] + ry - g.h, gz = c0[2] + rz - g.h;
            int wy = floor_div(gy, g.nc[1]), wz = floor_div(gz, g.nc[2]);
            int cy = gy - wy * g.nc[1], cz = gz - wz * g.nc[2];
            int wx = seg - 1;
            int lo = max(c0[0] - g.h, wx * g.nc[0]);
            int hi = min(c0[0] + e[0] + g.h - 1, (wx + 1) * g.nc[0] - 1);
            if (lo <= hi) {
                int cid_lo = (cz * g.nc[1] + cy) * g.nc[0] + (lo - wx * g.nc[0]);
                int cid_hi = (cz * g.nc[1] + cy) * g.nc[0] + (hi - wx * g.nc[0]);
                run.gstart = cell_start[cid_lo];
                run.count = cell_start[cid_hi + 1] - run.gstart;
                run.shift = (wx + 1) | ((wy + 1) << 2) | ((wz + 1) << 4);
                if (run.count > 0) len = ((run.gstart % A) + run.count + A - 1) / A * A;
            }
        }
        my_runs[r] = run;
        s_len[r] = len;
    }
    __syncthreads();
    // serial-in-chunks exclusive scan of s_len (max_runs <= a few hundred): warp 0, 32 at a time
    if (tid < 32) {
        int carry = A;  // slots [0, A) hold the dummy atom
        for (int base = 0; base < g.max_runs; base += 32) {
            int r = base + tid;
            int v = (r < g.max_runs) ? s_len[r] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            if (r < g.max_runs) s_base[r] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) s_total = carry;
    }
    __syncthreads();
    unsigned int tx_pos = 0, tx_lj = 0;
    int any_shift = 0;
    for (int r = tid; r < g.max_runs; r += blockDim.x) {
        Run run = my_runs[r];
        if (run.count > 0) {
            run.soff = s_base[r] + (run.gstart % A);
            my_runs[r].soff = run.soff;
            tx_pos += (unsigned int)run.count * (unsigned int)sizeof(typename VT<T>::T4);
            if (!uniform_lj) tx_lj += (unsigned int)s_len[r] * (unsigned int)sizeof(typename VT<T>::T2);
            if (run.shift != (1 | (1 << 2) | (1 << 4))) any_shift = 1;
        }
    }
    // block reduce tx / any_shift through shared atomics
    __shared__ unsigned int s_tx_pos, s_tx_lj;
    __shared__ int s_any;
    if (tid == 0) { s_tx_pos = 0; s_tx_lj = 0; s_any = 0; }
    __syncthreads();
    if (tx_pos) atomicAdd(&s_tx_pos, tx_pos);
    if (tx_lj) atomicAdd(&s_tx_lj, tx_lj);
    if (any_shift) atomicOr(&s_any, 1);
    __syncthreads();
    // halo cell table
    ushort2* my_hcs = hcs + (size_t)b * g.hcells;
    for (int hc = tid; hc < g.hcells; hc += blockDim.x) {
        int rx = hc % g.H[0], ry = (hc / g.H[0]) % g.H[1], rz = hc / (g.H[0] * g.H[1]);
        ushort2 se = make_ushort2(0, 0);
        if (rx < He[0] && ry < He[1] && rz < He[2]) {
            int gx = c0[0] + rx - g.h, gy = c0[1] + ry - g.h, gz = c0[2] + rz - g.h;
            int wx = floor_div(gx, g.nc[0]), wy = floor_div(gy, g.nc[1]), wz = floor_div(gz, g.nc[2]);
            int cid = ((gz - wz * g.nc[2]) * g.nc[1] + (gy - wy * g.nc[1])) * g.nc[0] + (gx - wx * g.nc[0]);
            int r = (rz * g.H[1] + ry) * 3 + (wx + 1);
            Run run = my_runs[r];
            int cs = cell_start[cid], ce = cell_start[cid + 1];
            int st = run.soff + (cs - run.gstart);
            int en = st + (ce - cs);
            st = min(st, 65535);
            en = min(en, 65535);
            se = make_ushort2((unsigned short)st, (unsigned short)en);
        }
        my_hcs[hc] = se;
    }
    __syncthreads();
    if (tid == 0) {
        IRow* my_rows = irows + (size_t)b * g.n_irows;
        int cum = 0;
        for (int q = 0; q < g.n_irows; q++) {
            int iy = q % g.b[1], iz = q / g.b[1];
            IRow row = {0, 0, 0, cum};
            if (iy < e[1] && iz < e[2]) {
                int cid0 = ((c0[2] + iz) * g.nc[1] + (c0[1] + iy)) * g.nc[0] + c0[0];
                row.slot_begin = cell_start[cid0];
                row.count = cell_start[cid0 + e[0]] - row.slot_begin;
                int hc = ((iz + g.h) * g.H[1] + (iy + g.h)) * g.H[0] + g.h;
                row.smem_begin = my_hcs[hc].x;
            }
            my_rows[q] = row;
            cum += row.count;
        }
        BrickHdr hd;
        hd.halo_count = s_total;
        hd.i_count = cum;
        hd.tx_pos = s_tx_pos;
        hd.tx_lj = s_tx_lj;
        hd.any_shift = s_any;
        hd.pad[0] = hd.pad[1] = hd.pad[2] = 0;
        hdrs[b] = hd;
        atomicMax(&ctl->max_halo, s_total);
        if (s_total > g.halo_cap) atomicOr(&ctl->overflow, 1);
    }
}

// ---- halo staging shared by the list builder and the force kernel -------------------------------
// Stages pos4 (and optionally lj2) runs of brick b into shared memory with bulk async copies, then
// converts the positions to the brick-local frame (origin = brick corner, periodic image applied) in
// double so that i-j differences carry no box-size rounding error.
// Phase 1: arm the mbarrier and issue the bulk copies (returns right after the issue; the copies are in flight).
template <typename T, bool WITH_LJ>
__device__ __forceinline__ void stage_halo_issue(const Geom<T>& g, const BrickHdr& hd, const Run* __restrict__ my_runs,
                                                 const typename VT<T>::T4* __restrict__ pos4,
                                                 const typename VT<T>::T2* __restrict__ lj2, typename VT<T>::T4* s_pos,
                                                 typename VT<T>::T2* s_lj, uint64_t* bar) {
    using T4 = typename VT<T>::T4;
    using T2 = typename VT<T>::T2;
    const int tid = threadIdx.x;
    const int A = g.align;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (tid < A) {
        s_pos[tid] = make4<T>((T)1.0e6, (T)1.0e6, (T)1.0e6, (T)0);  // dummy atom: far away, no charge
        if (WITH_LJ) s_lj[tid] = make2<T>((T)0, (T)0);
    }
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(bar, hd.tx_pos + (WITH_LJ ? hd.tx_lj : 0u));
    __syncthreads();
    for (int r = tid; r < g.max_runs; r += blockDim.x) {
        Run run = my_runs[r];
        if (run.count > 0) {
            bulk_g2s(&s_pos[run.soff], &pos4[run.gstart], (uint32_t)run.count * (uint32_t)sizeof(T4), bar);
            if (WITH_LJ) {
                int mis = run.gstart % A;
                int len = (mis + run.count + A - 1) / A * A;
                bulk_g2s(&s_lj[run.soff - mis], &lj2[run.gstart - mis], (uint32_t)len * (uint32_t)sizeof(T2), bar);
            }
        }
    }
}
// Phase 2: wait for the copies, then re-centre bricks that hold periodic images.
template <typename T, bool ALWAYS_LOCALIZE>
__device__ __forceinline__ void stage_halo_wait(const Geom<T>& g, int b, const BrickHdr& hd, const Run* __restrict__ my_runs,
                                                typename VT<T>::T4* s_pos, uint64_t* bar) {
    using T4 = typename VT<T>::T4;
    const int tid = threadIdx.x;
    mbar_wait(bar, 0);
    // Bricks whose halo holds no periodic image keep global coordinates: the pair loop only uses differences of
    // positions, which are exact in the same frame. Only bricks at the box boundary are re-centred.
    if (!ALWAYS_LOCALIZE && !hd.any_shift) {
        __syncthreads();
        return;
    }
    // The list builder works in a brick-local frame (ALWAYS_LOCALIZE: origin = brick corner, its row trimming needs
    // coordinates relative to the brick). The force kernel only needs every staged atom in the SAME frame as the owned
    // atoms, so it keeps global coordinates and moves just the runs that are periodic images by +-L (in double, one
    // rounding) - a fifth of a boundary brick's halo instead of all of it.
    double org[3] = {0.0, 0.0, 0.0};
    if (ALWAYS_LOCALIZE) {
        int B[3] = {b % g.nb[0], (b / g.nb[0]) % g.nb[1], b / (g.nb[0] * g.nb[1])};
#pragma unroll
        for (int d = 0; d < 3; d++) org[d] = (double)(B[d] * g.b[d]) * g.celld[d];
    }
    constexpr int NO_SHIFT = 1 | (1 << 2) | (1 << 4);
    // 8-lane groups, one run each (a run is ~40 atoms): four runs per warp instruction
    const int l8 = tid & 7, grp = tid >> 3, ngrp = blockDim.x >> 3;
    for (int r = grp; r < g.max_runs; r += ngrp) {
        Run run = my_runs[r];
        if (run.count <= 0) continue;
        if (!ALWAYS_LOCALIZE && run.shift == NO_SHIFT) continue;
        const double ox = (double)((run.shift & 3) - 1) * g.Ld[0] - org[0];
        const double oy = (double)(((run.shift >> 2) & 3) - 1) * g.Ld[1] - org[1];
        const double oz = (double)(((run.shift >> 4) & 3) - 1) * g.Ld[2] - org[2];
        for (int k = l8; k < run.count; k += 8) {
            T4 p = s_pos[run.soff + k];
            p.x = (T)((double)p.x + ox);
            p.y = (T)((double)p.y + oy);
            p.z = (T)((double)p.z + oz);
            s_pos[run.soff + k] = p;
        }
    }
    __syncthreads();
}
template <typename T, bool WITH_LJ, bool ALWAYS_LOCALIZE>
__device__ __forceinline__ void stage_halo(const Geom<T>& g, int b, const BrickHdr& hd, const Run* __restrict__ my_runs,
                                           const typename VT<T>::T4* __restrict__ pos4,
                                           const typename VT<T>::T2* __restrict__ lj2,
                                           typename VT<T>::T4* s_pos, typename VT<T>::T2* s_lj, uint64_t* bar) {
    stage_halo_issue<T, WITH_LJ>(g, hd, my_runs, pos4, lj2, s_pos, s_lj, bar);
    stage_halo_wait<T, ALWAYS_LOCALIZE>(g, b, hd, my_runs, s_pos, bar);
}

// Main-list entries are stored as halo index << LIST_SHIFT (= byte offset of a float4 position in shared memory): the
// force kernel saves a shift per entry. 16-bit entries therefore address at most LIST_MAX_HALO staged atoms per brick.
constexpr int LIST_SHIFT = 4;
constexpr int LIST_MAX_HALO = 65536 >> LIST_SHIFT;

// ---- R6: full-shell neighbour lists --------------------------------------------------------------
// One CTA per brick, one warp per owned atom. Entries are 16-bit halo indices written in the lane-
// swizzled order the force kernel reads (see force.cuh). Excluded pairs are dropped here; special
// (1-4) pairs go to a separate short list.
template <typename T, bool COUNT_ONLY, bool HAS_EX>
__global__ void __launch_bounds__(256)
    build_lists_kernel(Control* __restrict__ ctl, Geom<T> g, const BrickHdr* __restrict__ hdrs,
                       const Run* __restrict__ runs, const IRow* __restrict__ irows, const ushort2* __restrict__ hcs,
                       const typename VT<T>::T4* __restrict__ pos4, const int* __restrict__ orig,
                       const int* __restrict__ ex_ptr, const int* __restrict__ ex_idx, const int* __restrict__ sp_ptr,
                       const int* __restrict__ sp_idx, unsigned short* __restrict__ list,
                       unsigned short* __restrict__ slist, ushort2* __restrict__ counts, int brick0) {
    if (!ctl->rebuild) return;
    using T4 = typename VT<T>::T4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x + brick0;
    const BrickHdr hd = hdrs[b];
    if (hd.i_count == 0 || hd.halo_count > g.halo_cap) return;
    T4* s_pos = reinterpret_cast<T4*>(smem_raw);
    int* s_orig = reinterpret_cast<int*>(s_pos + g.halo_cap);
    ushort2* s_hcs = reinterpret_cast<ushort2*>(s_orig + g.halo_cap);
    IRow* s_rows = reinterpret_cast<IRow*>(s_hcs + ((g.hcells + 3) & ~3));
    __shared__ uint64_t s_bar;
    const Run* my_runs = runs + (size_t)b * g.max_runs;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    // far-away fill for every slot (covers alignment pads), then stage
    for (int k = tid; k < hd.halo_count; k += blockDim.x) {
        s_pos[k] = make4<T>((T)1.0e6, (T)1.0e6, (T)1.0e6, (T)0);
        s_orig[k] = -1;
    }
    for (int k = tid; k < g.hcells; k += blockDim.x) s_hcs[k] = hcs[(size_t)b * g.hcells + k];
    for (int k = tid; k < g.n_irows; k += blockDim.x) s_rows[k] = irows[(size_t)b * g.n_irows + k];
    __syncthreads();
    fence_proxy_async();
    stage_halo<T, false, true>(g, b, hd, my_runs, pos4, nullptr, s_pos, nullptr, &s_bar);
    for (int r = wid; r < g.max_runs; r += nw) {
        Run run = my_runs[r];
        for (int k = lane; k < run.count; k += 32) s_orig[run.soff + k] = orig[run.gstart + k];
    }
    __syncthreads();

    int my_max = 0;
    unsigned long long my_pairs = 0;
    for (int task = wid; task < hd.i_count; task += nw) {
        // locate the owned atom
        int q = 0;
        while (q + 1 < g.n_irows && s_rows[q + 1].cum <= task) q++;
        IRow row = s_rows[q];
        int k_in_row = task - row.cum;
        int slot = row.slot_begin + k_in_row;
        int si = row.smem_begin + k_in_row;
        int iy = q % g.b[1], iz = q / g.b[1];
        // which cell of the row holds it
        int ix = 0;
        {
            int hc0 = ((iz + g.h) * g.H[1] + (iy + g.h)) * g.H[0] + g.h;
            while (ix + 1 < g.b[0] && si >= (int)s_hcs[hc0 + ix].y) ix++;
        }
        T4 pi = s_pos[si];
        int oi = s_orig[si];
        // exclusion / special partner lists of atom oi into lanes
        int ex_a = ex_ptr ? ex_ptr[oi] : 0, ex_n = ex_ptr ? ex_ptr[oi + 1] - ex_a : 0;
        int sp_a = sp_ptr ? sp_ptr[oi] : 0, sp_n = sp_ptr ? sp_ptr[oi + 1] - sp_a : 0;
        int my_ex = (lane < ex_n) ? ex_idx[ex_a + lane] : -2;
        int my_sp = (lane < sp_n) ? sp_idx[sp_a + lane] : -2;
        int count = 0, scount = 0;
        unsigned short* my_list = list + (size_t)slot * g.stride;
        unsigned short* my_slist = slist + (size_t)slot * g.sstride;
        // Only the part of each halo row that can hold a neighbour is scanned: with dy, dz the distance from the
        // atom to the row's (y,z) cell slab, candidates need |dx| <= sqrt(r_list^2 - dy^2 - dz^2). The (2h+1)^2 rows
        // are handled 32 at a time: lane r trims row r, a warp scan turns the row lengths into offsets, and the
        // candidates of all rows are then walked as ONE flattened range (every lane busy; a 5-step shuffle search
        // maps a flattened index back to its row). Entry order = row-major, as a row-by-row scan would give.
        const T cyv = (T)g.celld[1], czv = (T)g.celld[2];
        const T inv_cx = g.inv_cell[0];
        const T rl2 = g.rlist2 * (T)1.0001;
        const int side = 2 * g.h + 1, nrows = side * side;
        for (int rb = 0; rb < nrows; rb += 32) {
            const int r = rb + lane;
            int ra = 0, rlen = 0;
            if (r < nrows) {
                const int rz = iz + r / side, ry = iy + r % side;
                const T zlo = (T)(rz - g.h) * czv;
                const T dzm = fmax(fmax(zlo - pi.z, pi.z - (zlo + czv)), (T)0);
                const T ylo = (T)(ry - g.h) * cyv;
                const T dym = fmax(fmax(ylo - pi.y, pi.y - (ylo + cyv)), (T)0);
                const T rem = rl2 - dym * dym - dzm * dzm;
                if (rem >= (T)0) {
                    const T wx = fsqrt(rem) + (T)1e-4;
                    int rx_lo = (int)ffloor((pi.x - wx) * inv_cx) + g.h;
                    int rx_hi = (int)ffloor((pi.x + wx) * inv_cx) + g.h;
                    rx_lo = max(rx_lo, ix);
                    rx_hi = min(rx_hi, ix + 2 * g.h);
                    const int hcrow = (rz * g.H[1] + ry) * g.H[0];
                    ra = s_hcs[hcrow + rx_lo].x;
                    rlen = max((int)s_hcs[hcrow + rx_hi].y - ra, 0);
                }
            }
            int inc = rlen;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const int roff = inc - rlen;  // exclusive offset of this lane's row in the flattened range
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            for (int k0 = 0; k0 < total; k0 += 32) {
                const int k = k0 + lane;
                int lo = 0;  // last row whose offset is <= k (empty rows share their successor's offset)
#pragma unroll
                for (int st = 16; st > 0; st >>= 1) {
                    int v = __shfl_sync(0xffffffffu, roff, lo + st);
                    if (v <= k) lo += st;
                }
                const int c = __shfl_sync(0xffffffffu, ra, lo) + (k - __shfl_sync(0xffffffffu, roff, lo));
                bool in = false, special = false;
                if (k < total && c != si) {
                    T4 pj = s_pos[c];
                    T dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    T d2 = dx * dx + dy * dy + dz * dz;
                    in = d2 <= g.rlist2;
                }
                int oj = (HAS_EX && in) ? s_orig[c] : -1;
                // exclusions (warp-uniform loops over the partner lists)
                if (HAS_EX && ex_n > 0) {
                    int nn = min(ex_n, 32);
                    for (int kk = 0; kk < nn; kk++) {
                        int v = __shfl_sync(0xffffffffu, my_ex, kk);
                        if (v == oj) in = false;
                    }
                    for (int kk = 32; kk < ex_n; kk++)
                        if (ex_idx[ex_a + kk] == oj) in = false;
                }
                if (HAS_EX && sp_n > 0) {
                    int nn = min(sp_n, 32);
                    for (int kk = 0; kk < nn; kk++) {
                        int v = __shfl_sync(0xffffffffu, my_sp, kk);
                        if (in && v == oj) special = true;
                    }
                    for (int kk = 32; kk < sp_n; kk++)
                        if (in && sp_idx[sp_a + kk] == oj) special = true;
                }
                bool main_hit = in && !special;
                bool spec_hit = in && special;
                unsigned int mb_ = __ballot_sync(0xffffffffu, main_hit);
                unsigned int sb_ = __ballot_sync(0xffffffffu, spec_hit);
                unsigned int lt = (1u << lane) - 1u;
                if (!COUNT_ONLY) {
                    if (main_hit) {
                        int m = count + __popc(mb_ & lt);
                        if (m < g.stride) {
                            int phys = (m & ~31) + ((m & 7) << 2) + ((m & 31) >> 3);
                            my_list[phys] = (unsigned short)(c << LIST_SHIFT);
                        }
                    }
                    if (spec_hit) {
                        int m = scount + __popc(sb_ & lt);
                        if (m < g.sstride) my_slist[m] = (unsigned short)c;
                    }
                }
                count += __popc(mb_);
                scount += __popc(sb_);
            }
        }
        if (!COUNT_ONLY) {
            // pad the last group of 32 with the dummy atom (halo slot 0)
            int padded = min((count + 31) & ~31, g.stride);
            for (int m = count + lane; m < padded; m += 32) {
                int phys = (m & ~31) + ((m & 7) << 2) + ((m & 31) >> 3);
                my_list[phys] = 0;
            }
            if (lane == 0) {
                counts[slot] = make_ushort2((unsigned short)min(count, g.stride), (unsigned short)min(scount, g.sstride));
                if (count > g.stride) atomicOr(&ctl->overflow, 2);
                if (scount > g.sstride) atomicOr(&ctl->overflow, 4);
            }
        }
        my_max = max(my_max, count);
        if (lane == 0) {
            my_pairs += (unsigned long long)(count + scount);
            atomicMax(&ctl->max_special, scount);
        }
    }
    if (lane == 0) {
        atomicMax(&ctl->max_neighbors, my_max);
        atomicAdd(&ctl->n_pairs, my_pairs);
    }
}

// ---- R7: finish ------------------------------------------------------------------------------
__global__ void rebuild_finish_kernel(Control* ctl) {
    if (!ctl->rebuild) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (ctl->disp) ctl->violations++;
        ctl->disp = 0;
        ctl->rebuild = 0;
        ctl->prune = 1;  // fresh outer lists: derive the inner lists from them
        ctl->n_rebuilds++;
        if (ctl->max_disp2_bits > ctl->call_max_disp2_bits) ctl->call_max_disp2_bits = ctl->max_disp2_bits;
        ctl->max_disp2_bits = 0;
    }
}
// ---- dual-list pruning -------------------------------------------------------------------------------
// The lists built above hold every pair within r_list (outer radius). The force kernel walks a shorter inner
// list: the pairs within r_inner = max r_cut + inner skin at the last prune, refreshed whenever an atom moved more
// than half the inner skin (cheap: no cell search, the outer list is the candidate set). Same 16-bit halo indices,
// same lane-swizzled layout, order preserved. One CTA per brick, 8 lanes per owned atom.
template <typename T>
__global__ void __launch_bounds__(256)
    prune_lists_kernel(const Control* __restrict__ ctl, Geom<T> g, const BrickHdr* __restrict__ hdrs, const Run* __restrict__ runs,
                       const IRow* __restrict__ irows, const typename VT<T>::T4* __restrict__ pos4,
                       const unsigned short* __restrict__ olist, const ushort2* __restrict__ ocounts,
                       unsigned short* __restrict__ ilist, ushort2* __restrict__ icounts,
                       typename VT<T>::T4* __restrict__ xprune4, int brick0) {
    if (!ctl->prune) return;
    using T4 = typename VT<T>::T4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x + brick0;
    const BrickHdr hd = hdrs[b];
    if (hd.i_count == 0 || hd.halo_count > g.halo_cap) return;
    T4* s_pos = reinterpret_cast<T4*>(smem_raw);
    __shared__ uint64_t s_bar;
    __shared__ IRow s_rows[64];
    const int tid = threadIdx.x;
    const Run* my_runs = runs + (size_t)b * g.max_runs;
    for (int k = tid; k < g.n_irows; k += blockDim.x) s_rows[k] = irows[(size_t)b * g.n_irows + k];
    stage_halo<T, false, false>(g, b, hd, my_runs, pos4, nullptr, s_pos, nullptr, &s_bar);
    const int sub = tid >> 3, l = tid & 7;
    const unsigned int sub_mask = 0xffu << (8 * (sub & 3));
    const unsigned int lt = (1u << l) - 1u;
    for (int task = sub; task < hd.i_count; task += 32) {
        int q = 0;
        while (q + 1 < g.n_irows && s_rows[q + 1].cum <= task) q++;
        const IRow row = s_rows[q];
        const int slot = row.slot_begin + (task - row.cum);
        const int si = row.smem_begin + (task - row.cum);
        const T4 pi = s_pos[si];
        const ushort2 cnt = ocounts[slot];
        const int n_groups = ((int)cnt.x + 31) >> 5;
        const unsigned short* lp = olist + (size_t)slot * g.stride;
        unsigned short* op = ilist + (size_t)slot * g.stride;
        int out = 0;
        for (int gi = 0; gi < n_groups; gi++) {
            const uint2 w = reinterpret_cast<const uint2*>(lp + gi * 32)[l];
            const int j[4] = {(int)(w.x & 0xffffu), (int)(w.x >> 16), (int)(w.y & 0xffffu), (int)(w.y >> 16)};
#pragma unroll
            for (int e = 0; e < 4; e++) {  // logical entry index inside the group = l + 8 e
                const T4 pj = s_pos[j[e] >> LIST_SHIFT];
                const T dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const bool in = (dx * dx + dy * dy + dz * dz) <= g.rinner2;
                const unsigned int bal = (__ballot_sync(sub_mask, in) >> (8 * (sub & 3))) & 0xffu;
                if (in) {
                    const int m = out + __popc(bal & lt);
                    op[(m & ~31) + ((m & 7) << 2) + ((m & 31) >> 3)] = (unsigned short)j[e];
                }
                out += __popc(bal);
            }
        }
        const int padded = (out + 31) & ~31;
        for (int m = out + l; m < padded; m += 8) op[(m & ~31) + ((m & 7) << 2) + ((m & 31) >> 3)] = 0;
        if (l == 0) {
            icounts[slot] = make_ushort2((unsigned short)out, cnt.y);
            xprune4[slot] = pos4[slot];
        }
    }
}
__global__ void prune_finish_kernel(Control* ctl) {
    if (!ctl->prune) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->prune = 0;
        ctl->n_prunes++;
    }
}

__global__ void rebuild_begin_kernel(Control* ctl) {
    if (!ctl->rebuild) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->n_pairs = 0;
        ctl->max_neighbors = 0;
        ctl->max_halo = 0;
        ctl->max_special = 0;
    }
}

}  // namespace mb
