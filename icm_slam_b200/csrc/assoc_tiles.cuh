// assoc_tiles.cuh -- k_assoc_tiles: data association observation by observation + construction of the run records.
//
// Runs over the tiles the steady-state kernel (runs.cuh k_runs) could not certify: every tile on the first sweep of a map
// chain or after landmarks were renumbered, afterwards only tiles with a run outside its landmark's proven radius or with far
// observations.  For every scan of the tile (sensors.py:146-154 restated per SURVEY.md App. A):
//   * the tile's contiguous slice of the observation records (bx, by) and last sweep's labels are staged in shared memory
//     with TMA bulk copies (cp.async.bulk + mbarrier); a tile whose observations exceed the budget goes in chunks of whole
//     scans;
//   * phase A -- LANES OVER CONSECUTIVE OBSERVATIONS (four per lane in flight): projection with the sweep's input pose
//     (tras_rot_z, ICM_SLAM.py:465-480, numpy's exact operation order); last sweep's label of the same observation is
//     accepted iff the point lies inside the landmark's proven-nearest radius (tail.cuh hint_radius2), else the landmark grid
//     (fastgrid.cuh) gives the nearest landmark by squared distance with the rooted values formed only inside a 2^-50 tie
//     window, and the gate s > thr2_hi <=> sqrt(s) > dist_thr: bit-exact cdist + argmin + gate (ICM_SLAM.py:169-172);
//   * pass 1 -- THREAD PAIR PER SCAN: each maximal group of beams with the same label becomes a run (sum of the body-frame
//     points, beam count, largest distance of a beam from the centroid as the bound rho of runs.cuh);
//   * the runs are packed scan by scan into 32-record chunks (a scan never straddles a chunk unless it has more than 32
//     runs) and written to the tile's region of the record array;
//   * the chunks are then processed by the SAME code as the steady state (process_chunk), so moments and statistics do not
//     depend on which of the two kernels handled a scan.
#pragma once
#include "common.cuh"
#include "assoc.cuh"
#include "fastgrid.cuh"
#include "tail.cuh"
#include "runs.cuh"

#define AT_THREADS (2 * RT_TILE)   // a thread pair per scan of the tile
#define AT_WARPS (AT_THREADS / 32)
#define AT_U 4                     // observations per lane in flight in phase A (pend bitmask: 64 / AT_U iterations per warp)
#define AT_OBS_BYTES 25            // staged bytes per observation: (bx, by) 16, label 4, run length 2, rho code 2, scan 1

struct AssocParams {
    int first_halo;                       // the first scan of tile 0 is the halo scan of a time segment (moments only)
    const double2* bxy;                   // body-frame observations (bx, by), CSR order
    DevCfg cfg;
    double thr2_hi;                       // largest s with sqrt_rn(s) <= dist_thr
    const DevState* st;
    const FGeom* geom;
    const int* cell_start; const double2* gpts; const int* gidx;
    const int* remap;                     // last sweep's label -> this map's label (hints)
    int hints;                            // c[] holds last sweep's labels for this map chain
    int* c;                               // labels per observation (out)
    int obs_cap;                          // shared-memory capacity in observations
    int skip_hints;                       // debug (ICMSLAM_HINTS=0)
    RunParams R;
};

struct __align__(16) AssocSmem {
    double2 pp[RT_TILE];           // projection origin of each scan (self.x0 for scan 0)
    double2 rsc[RT_TILE];          // (sin, cos) of (projection heading - pi/2)
    int off[RT_TILE + 4];          // off[t] of the tile's scans
    int rcnt[RT_TILE];             // runs of each scan
    int rstart[RT_TILE];           // slot of the scan's first run, relative to the tile's region
    int gap[RT_TILE + 2][2];       // padding slots [begin, end)
    int ngap, pos;
    unsigned long long mbar;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(AT_THREADS, 2)
k_assoc_tiles(const AssocParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AssocSmem& S = *reinterpret_cast<AssocSmem*>(smem_raw);
    double2* sb = reinterpret_cast<double2*>(smem_raw + sizeof(AssocSmem));             // staged observations; later run sums
    int* sbk_raw = reinterpret_cast<int*>(sb + p.obs_cap);                               // hints in (TMA), labels out; +4 ints of alignment slack
    unsigned short* srn = reinterpret_cast<unsigned short*>(sbk_raw + p.obs_cap + 4);    // run length at run heads (phase A: cell entry count)
    unsigned short* srho = srn + p.obs_cap;                                              // rho code at run heads
    unsigned char* slt = reinterpret_cast<unsigned char*>(srho + p.obs_cap);             // scan (relative to the tile) of each observation

    const RunParams& R = p.R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mb = smem_u32(&S.mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    FGrid G;
    G.g = *p.geom;
    G.cell_start = p.cell_start; G.pts = p.gpts; G.idx = p.gidx;
    const int lsearch = p.st->lsearch;
    const bool have_map = lsearch > 0;
    const bool ident = R.ts->remap_identity != 0;
    const bool use_hints = p.hints != 0 && have_map && !p.skip_hints;
    const int epoch = R.ts->epoch;
    const int n_dirty = R.ts->n_dirty;
    const int q = tid >> 1, sub = tid & 1;          // scan of the tile worked by this thread pair, and the thread's half of it

    for (int di = blockIdx.x; di < n_dirty; di += gridDim.x) {
        const int tile = R.dirty_list[di];
        const int tb = R.t_start + tile * RT_TILE;
        const int nsc = min(RT_TILE, R.t_hi - tb);
        const bool commit_all = R.tile_flag[tile] == 2;
        const bool halo_tile = p.first_halo && tile == 0;
        const int64_t base = run_tile_base(R.off, R.t_start, tile);
        // ---- tile loads -------------------------------------------------------------------------------------------
        for (int li = tid; li <= nsc; li += AT_THREADS) S.off[li] = R.off[tb + li];
        if (tid < nsc) {
            const double4 pq = ldg_ppar(R.ppar + tb + tid);
            S.pp[tid] = make_double2(pq.x, pq.y); S.rsc[tid] = make_double2(pq.z, pq.w);
        }
        if (tid == 0) { S.pos = 0; S.ngap = 0; }
        __syncthreads();
        // ---- chunks of whole scans whose observations fit the shared-memory budget (normally one) -----------------
        for (int c_lo = 0; c_lo < nsc;) {
            int c_hi = nsc - 1;
            if (S.off[nsc] - S.off[c_lo] > p.obs_cap) {
                int lo = c_lo, hi = nsc - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (S.off[mid + 1] - S.off[c_lo] <= p.obs_cap) lo = mid; else hi = mid - 1;
                }
                c_hi = lo;
            }
            const int co = S.off[c_lo], ce = S.off[c_hi + 1];
            const int co4 = co & ~3;                         // 16-byte aligned start of the labels' slice
            int* sbk = sbk_raw + (co - co4);                 // label of each observation (-1 far), same slot as its hint
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)(ce - co) * 16u;
                const uint32_t hbytes = (use_hints && ce > co) ? (uint32_t)((ce - co4 + 3) & ~3) * 4u : 0u;
                if (bytes + hbytes > 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes + hbytes) : "memory");
                    if (bytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                            ::"r"(smem_u32(sb)), "l"(p.bxy + co), "r"(bytes), "r"(mb) : "memory");
                    // last sweep's labels of the same observations (the hints) land in the label array itself
                    if (hbytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                             ::"r"(smem_u32(sbk_raw)), "l"(p.c + co4), "r"(hbytes), "r"(mb) : "memory");
                } else {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
                }
            }
            // the pair's scan and this thread's half of it (indices relative to the chunk)
            const bool mine = q >= c_lo && q <= c_hi;
            int o = 0, e = 0;
            if (mine) {
                const int so = S.off[q] - co, se = S.off[q + 1] - co;
                const int h1 = (se - so + 1) >> 1;
                o = min(so + sub * h1, se);
                e = min(o + h1, se);
            }
            for (int i = o; i < e; ++i) slt[i] = (unsigned char)q;     // (while the bulk copies are in flight)
            mbar_wait(mb, parity);
            parity ^= 1u;
            __syncthreads();
            // ---- phase A: lanes over consecutive observations (adjacent lanes = adjacent beams), four per lane in flight --
            //   A0 (hints): last sweep's label of the same observation, carried through the filter's renumbering, is
            //       accepted when the observation lies inside the landmark's proven-nearest radius (tail.cuh hint_radius2):
            //       one staged value and one gather instead of the grid search.  Lanes it cannot settle stay pending for
            //   A1: project, locate the grid cell, fetch its entry range     (first level of gathers)
            //   A2: fetch the candidates, pick the nearest, gate, label      (second level of gathers)
            {
                const int m = ce - co;
                const int per = ((((m + AT_WARPS - 1) / AT_WARPS) + AT_U * 32 - 1) / (AT_U * 32)) * (AT_U * 32);
                const int wa = warp * per, wb = min(wa + per, m);
                unsigned long long pend = 0ull;
                int slot = 0;
                for (int b0 = wa; b0 < wb; b0 += AT_U * 32, slot += AT_U) {
                    int h_[AT_U];
                    double wx_[AT_U], wy_[AT_U];
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        const int i = b0 + u * 32 + lane;
                        h_[u] = -1; wx_[u] = 0.0; wy_[u] = 0.0;
                        if (i < wb && use_hints) {
                            h_[u] = sbk[i];
                            if (h_[u] >= p.cfg.L) h_[u] = -1;          // (labels of scans this handle does not own are not maintained)
                            const double2 bq = sb[i];
                            const int lt = slt[i];
                            const double2 pq = S.pp[lt], rc = S.rsc[lt];
                            wx_[u] = add_rn(__fma_rn(bq.y, -rc.x, mul_rn(bq.x, rc.y)), pq.x);
                            wy_[u] = add_rn(__fma_rn(bq.y, rc.y, mul_rn(bq.x, rc.x)), pq.y);
                        }
                    }
                    if (!ident) {
#pragma unroll
                        for (int u = 0; u < AT_U; ++u) if (h_[u] >= 0) h_[u] = __ldg(p.remap + h_[u]);
                    }
                    double2 q_[AT_U];
                    double r2_[AT_U];
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        q_[u] = make_double2(0.0, 0.0); r2_[u] = -1.0;
                        if (h_[u] >= 0 && h_[u] < lsearch) {
                            const double2* rp = reinterpret_cast<const double2*>(R.lmrec + h_[u]);
                            q_[u] = __ldg(rp); r2_[u] = __ldg(rp + 1).x;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        const int i = b0 + u * 32 + lane;
                        if (i < wb) {
                            const double d2 = dist2_rn(q_[u].x - wx_[u], q_[u].y - wy_[u]);
                            if (d2 <= r2_[u]) {
                                sbk[i] = h_[u];
                                if (!ident && !(halo_tile && slt[i] == 0)) p.c[co + i] = h_[u];    // (with the identity renumbering c[] already holds it)
                            } else pend |= 1ull << (slot + u);
                        }
                    }
                }
                const bool any_pend = __any_sync(FULLMASK, pend != 0ull);
                slot = 0;
                for (int b0 = wa; b0 < wb && any_pend; b0 += AT_U * 32, slot += AT_U) {
                    int s_[AT_U], e_[AT_U];
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        const int i = b0 + u * 32 + lane;
                        s_[u] = 0; e_[u] = 0;
                        if (i < wb && have_map && ((pend >> (slot + u)) & 1ull)) {
                            const double2 bq = sb[i];
                            const int lt = slt[i];
                            const double2 pq = S.pp[lt], rc = S.rsc[lt];
                            const double wx = add_rn(__fma_rn(bq.y, -rc.x, mul_rn(bq.x, rc.y)), pq.x);
                            const double wy = add_rn(__fma_rn(bq.y, rc.y, mul_rn(bq.x, rc.x)), pq.y);
                            const int cc = fgrid_cell(G.g, wx, wy);
                            s_[u] = __ldg(G.cell_start + cc); e_[u] = __ldg(G.cell_start + cc + 1);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        const int i = b0 + u * 32 + lane;
                        if (i < wb && ((pend >> (slot + u)) & 1ull)) { sbk[i] = s_[u]; srn[i] = (unsigned short)min(e_[u] - s_[u], 65535); }
                    }
                }
                slot = 0;
                for (int b0 = wa; b0 < wb && any_pend; b0 += AT_U * 32, slot += AT_U) {
                    double wx_[AT_U], wy_[AT_U];
                    double2 p_[AT_U];
                    int id_[AT_U], s_[AT_U], n_[AT_U];
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        const int i = b0 + u * 32 + lane;
                        s_[u] = 0; n_[u] = 0; id_[u] = -1; p_[u] = make_double2(0.0, 0.0); wx_[u] = 0.0; wy_[u] = 0.0;
                        if (i < wb && ((pend >> (slot + u)) & 1ull)) {
                            s_[u] = sbk[i]; n_[u] = srn[i];
                            const double2 bq = sb[i];
                            const int lt = slt[i];
                            const double2 pq = S.pp[lt], rc = S.rsc[lt];
                            // tras_rot_z: numpy's matmul order, acc = a0*b0; acc = fma(a1, b1, acc); + translation
                            wx_[u] = add_rn(__fma_rn(bq.y, -rc.x, mul_rn(bq.x, rc.y)), pq.x);
                            wy_[u] = add_rn(__fma_rn(bq.y, rc.y, mul_rn(bq.x, rc.x)), pq.y);
                            if (n_[u] > 0) { p_[u] = __ldg(G.pts + s_[u]); id_[u] = __ldg(G.idx + s_[u]); }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < AT_U; ++u) {
                        const int i = b0 + u * 32 + lane;
                        if (i < wb && ((pend >> (slot + u)) & 1ull)) {
                            double best;
                            int bid;
                            const int bk = fgrid_scan_pre(G, wx_[u], wy_[u], s_[u], n_[u], p_[u], id_[u], best, bid);
                            const bool far = bk < 0 || best > p.thr2_hi;      // amin > dist_thr (ICM_SLAM.py:172)
                            sbk[i] = far ? -1 : bid;                          // the label (index in the previous map)
                            if (!(halo_tile && slt[i] == 0)) p.c[co + i] = far ? -1 : bid;    // the halo scan is not owned
                        }
                    }
                }
            }
            __syncthreads();     // the labels of the whole chunk are visible to the scan threads
            // ---- pass 1 (thread pair per scan): runs of equal labels -> in-place run records --------------------------
            int nr = 0;
            if (o < e) {
                // (the next observation is fetched before the current one is consumed)
                int run_start = o, bk = sbk[o];
                double2 b = sb[o], bfirst = b;
                double Sbx = 0.0, Sby = 0.0;
                for (int i = o; i < e; ++i) {
                    const int inx = min(i + 1, e - 1);
                    const double2 bn = sb[inx];
                    const int bkn = sbk[inx];
                    Sbx += b.x; Sby += b.y;
                    if (i + 1 == e || bkn != bk) {
                        // rho = max |b_i - centroid| over the run's beams (second walk: the run's other beams are still staged)
                        const double inv = 1.0 / (double)(i + 1 - run_start), cx = Sbx * inv, cy = Sby * inv;
                        double r2m = fma(bfirst.x - cx, bfirst.x - cx, (bfirst.y - cy) * (bfirst.y - cy));
                        for (int j = run_start + 1; j <= i; ++j) {
                            const double2 bj = sb[j];
                            r2m = fmax(r2m, fma(bj.x - cx, bj.x - cx, (bj.y - cy) * (bj.y - cy)));
                        }
                        const double rho = sqrt(r2m) * (1.0 + 1e-6) + 1e-9;
                        sb[run_start] = make_double2(Sbx, Sby);
                        srn[run_start] = (unsigned short)(i + 1 - run_start);
                        srho[run_start] = (unsigned short)min(__double2int_ru(rho * RT_RHO_UNIT), 65535);
                        ++nr;
                        run_start = i + 1; Sbx = 0.0; Sby = 0.0; bfirst = bn;
                    }
                    b = bn; bk = bkn;
                }
            }
            const int nr_other = __shfl_xor_sync(FULLMASK, nr, 1);
            const int rtot = nr + nr_other, kbase = sub ? nr_other : 0;
            if (mine && sub == 0) S.rcnt[q] = rtot;
            __syncthreads();
            // ---- packing: scans in time order into chunks of 32 slots -------------------------------------------------
            if (tid == 0) {
                int pos = S.pos, ng = S.ngap;
                for (int lt = c_lo; lt <= c_hi; ++lt) {
                    const int r = S.rcnt[lt];
                    const int fill = pos & 31;
                    if (r > 0 && fill != 0 && r > 32 - fill) {       // does not fit (a scan of more than 32 runs starts a chunk)
                        S.gap[ng][0] = pos; S.gap[ng][1] = pos + 32 - fill; ++ng;
                        pos += 32 - fill;
                    }
                    S.rstart[lt] = pos;
                    pos += r;
                }
                S.pos = pos; S.ngap = ng;
            }
            __syncthreads();
            // ---- emission: the pair's runs to the tile's region ---------------------------------------------------------
            if (o < e) {
                const int s0 = S.rstart[q];
                const bool lng = rtot > 32;
                const int npieces = (rtot + 31) >> 5;
                int k = kbase;
                for (int i = o; i < e; ++k) {
                    const int n = srn[i];
                    const int piece = lng ? (k >> 5) : 0, kk = lng ? (k & 31) : k;
                    const int plen = lng ? min(32, rtot - 32 * piece) : rtot;
                    const unsigned rc = srho[i];
                    const float rho = rc >= 65535u ? INFINITY : (float)rc * (float)(1.0 / RT_RHO_UNIT);
                    const int flags = (kk == 0 ? RF_LEADER : 0) | (lng ? RF_LONG : 0) | ((halo_tile && q == 0) ? RF_HALO : 0);
                    RunRec* rp = R.rec + base + s0 + k;
                    *reinterpret_cast<double2*>(rp) = sb[i];
                    *(reinterpret_cast<int4*>(rp) + 1) = make_int4(sbk[i], __float_as_int(rho), n | (q << 16) | ((plen - 1 - kk) << 24),
                                                                   piece | (flags << 8) | (npieces << 16));
                    i += n;
                }
            }
            c_lo = c_hi + 1;
            if (c_lo < nsc) __syncthreads();   // the next chunk overwrites the staging buffers
        }
        __syncthreads();
        // ---- padding slots, tile header --------------------------------------------------------------------------------
        {
            const int pos = S.pos, fill = pos & 31;
            const int nch = (pos + 31) >> 5;
            const int ng = S.ngap;
            for (int g = 0; g <= ng; ++g) {
                const int gb = g < ng ? S.gap[g][0] : pos, ge = g < ng ? S.gap[g][1] : (fill ? pos + 32 - fill : pos);
                for (int s = gb + tid; s < ge; s += AT_THREADS) {
                    RunRec* rp = R.rec + base + s;
                    *reinterpret_cast<double2*>(rp) = make_double2(0.0, 0.0);
                    *(reinterpret_cast<int4*>(rp) + 1) = make_int4(RUN_PAD, 0, 0, 0);
                }
            }
            if (tid == 0) { R.tile_nchunks[tile] = nch; R.tile_epoch[tile] = epoch; }
            __syncthreads();     // the block's records are visible to its warps
            for (int c = warp; c < nch; c += AT_WARPS)
                process_chunk<false>(R, R.rec + base + (int64_t)c * 32, (base >> 5) + c, tb, tile, commit_all);
        }
        __syncthreads();
        if (tid < nsc) R.scan_dirty[tb + tid] = 0;
        if (tid == 0) R.tile_flag[tile] = 0;
    }
}

static size_t assoc_smem_bytes(int obs_cap)   // obs_cap is even
{
    return sizeof(AssocSmem) + (size_t)obs_cap * AT_OBS_BYTES + 16 + 64;
}
