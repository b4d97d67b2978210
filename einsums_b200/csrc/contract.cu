// contract.cu — host launcher of the contraction engine behind eb200_contract and every BLAS-like
// entry point. Replaces the device einsum_generic_algorithm launcher
// (reference: libs/Einsums/TensorAlgebra/include/Einsums/TensorAlgebra/Backends/GPUTensorAlgebra.hpp:200-342).
#include "contract.cuh"
#include "contract_kernels.cuh"
#include "gemm_tma.cuh"
#include "contract_tma.cuh"
#include "gemm_tma_z.cuh"
#include "launch_api.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include <cstdlib>

namespace eb200 {
namespace {

int grid_for(int64_t total, int threads, int per_sm = 8) {
    int64_t need = (total + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * per_sm;
    return (int)std::max<int64_t>(1, std::min(need, cap));
}

template <typename K> int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) EB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return EB200_OK;
}

// stream-ordered scratch for split partials: no synchronisation, freed after the finalize kernel (scratch_alloc, runtime.cu:
// a PRIVATE memory pool with a bounded release threshold; the device's default pool is left alone)
int ws_alloc(void **p, size_t bytes, cudaStream_t s) { return scratch_alloc(p, bytes, s); }

template <typename T> int launch_scale(cudaStream_t s, const Grp &t, void *C, const double beta[2], eb200_contract_desc &d) {
    ScaleArgs a;
    a.t = t;
    a.C = C;
    a.beta[0] = beta[0];
    a.beta[1] = beta[1];
    if (t.tot == 0) return EB200_OK;
    { int rc_ = k_launch_scale(Scalar<T>::dtype, s, a, grid_for(t.tot, 256)); if (rc_) return rc_; }
    d.launches++;
    return EB200_OK;
}

template <typename T>
int launch_finalize(cudaStream_t s, const Grp &gb, const Grp &gm, const Grp &gn, const void *ws, void *C, const double alpha[2],
                    const double beta[2], int splits, eb200_contract_desc &d) {
    FinalizeArgs a;
    a.gb = gb; a.gm = gm; a.gn = gn;
    a.ws = ws; a.C = C;
    a.alpha[0] = alpha[0]; a.alpha[1] = alpha[1];
    a.beta[0] = beta[0];   a.beta[1] = beta[1];
    a.splits = splits;
    int64_t total = gb.tot * gm.tot * gn.tot;
    { int rc_ = k_launch_finalize(Scalar<T>::dtype, s, a, grid_for(total, 256)); if (rc_) return rc_; }
    d.launches++;
    return EB200_OK;
}

// ---------------------------------------------------------------- tile kernels
struct TileChoice { int bm, bn, bk, per_sm; };

inline bool even(int64_t x) { return (x & 1) == 0; }
// every stride of `role` in group g is even, except index `skip` (-1: none)
bool strides_even(const Grp &g, int role, int skip) {
    for (int i = 0; i < g.n; i++)
        if (i != skip && !even(g.s[role][i])) return false;
    return true;
}
// operand `role` can be fetched as aligned pairs along the inner index of `fast`
bool operand_pairs_ok(const Plan &p, int role, const Grp &fast, const Grp &other, const void *base) {
    if (fast.n == 0 || fast.s[role][0] != 1 || !even(fast.ext[0])) return false;
    if (!strides_even(fast, role, 0) || !strides_even(other, role, -1) || !strides_even(p.b, role, -1)) return false;
    return ((uintptr_t)base % 16) == 0;
}

int g_mma_config = 0; // 0 = auto; 1 = 128x128 (1 CTA/SM); 2 = 128x64 (2 CTA/SM). Tuning hook (EB200_MMA_CONFIG).
std::once_flag g_cfg_once;

// ---------------------------------------------------------------- TMA fast path (plain FP64 GEMM)
// tuning / testing knobs (eb200_set_option, defaults from the environment): tma = 0 disables both TMA kernels; tma_generic = 0
// disables the multi-index one, 2 takes it whenever it is legal (ignores the padding-waste rule); tma_tile = 0 / 1 forces 128x64 / 80x80

// 2-D map over an operand: dim0 = the unit-stride direction, dim1 = the other one (ld elements apart)
bool make_operand_map(CUtensorMap *map, const double *base, int64_t inner, int64_t outer, int64_t ld, uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int env_int_early(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---- tile shape / split-K choice shared by both TMA kernels. mb, nb: 16-row block counts of the (padded) composite M and N.
struct TmaChoice { int cfg; int64_t tm, tn; int splits; };
TmaChoice choose_tma_tile(int64_t mb, int64_t nb, int64_t batches, int64_t KT, int forced_split) {
    const int forced_cfg = option(OPT_TMA_TILE);
    const int64_t tm0 = (mb + 7) / 8, tn0 = (nb + 3) / 4, tm1 = (mb + 4) / 5, tn1 = (nb + 4) / 5;
    const double  real = (double)mb * nb, area0 = (double)tm0 * 8 * tn0 * 4, area1 = (double)tm1 * 5 * tn1 * 5;
    TmaChoice     c;
    // 128x64 unless it pads the output by more than 15 % and 80x80 tiles cover it clearly better (Wmnij: 400x400 -> 512x448 vs 400x400)
    c.cfg = (area0 > 1.15 * real && area1 * 1.08 < area0) ? 1 : 0;
    if (forced_cfg == 0 || forced_cfg == 1) c.cfg = forced_cfg;
    c.tm = c.cfg ? tm1 : tm0;
    c.tn = c.cfg ? tn1 : tn0;
    const int64_t tiles = c.tm * c.tn * std::max<int64_t>(1, batches);
    int64_t       splits = 1;
    if (forced_split > 0) splits = forced_split;
    else {
        // Split-K against wave quantisation: with `target` CTAs resident at a time, t tiles x s splits run in ceil(t s / target) waves.
        // 175 tiles on 296 slots (one 800 x 1600 x 80000 product of the K-split sharding) keep 59 % of the machine busy; 5 splits, 98.5 %.
        // Cost model in k-steps: a wave lasts KT / s steps plus, for s > 1, the partial-tile store and load (~1.5 steps) and the last
        // CTA's serial sum over s partials (~0.75 s steps). 3000^3 (KT = 188) stays unsplit, 1600 x 1600 x 160000 takes 10 splits
        // (25.5 -> 34.2 TFLOP/s), 800 x 1600 x 80000 takes 5 (19 -> 31).
        const int64_t target = (int64_t)sm_count() * (c.cfg ? 3 : 2);
        if (tiles * 2 <= target) splits = std::min<int64_t>(target / tiles, std::max<int64_t>(1, KT / 8));
        else if (tiles < 6 * target) {
            const int64_t smax = std::min<int64_t>(32, KT / 16);
            double        best = 0.0;
            for (int64_t sp = 1; sp <= smax; sp++) {
                const int64_t waves = (tiles * sp + target - 1) / target;
                const double  cost = (double)waves * ((double)KT / (double)sp + (sp > 1 ? 1.5 + 0.75 * (double)sp : 0.0));
                if (sp == 1 || cost < best * 0.97) { best = cost; splits = sp; } // a split has to win by 3 % to be taken
            }
        }
    }
    c.splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(splits, KT), 65535));
    return c;
}

// returns 1 when the TMA kernel ran, 0 when the shape does not qualify, <0 on error
int try_run_tma(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    if (!option(OPT_TMA) || (d.force_kernel != 0 && d.force_kernel != EB200_KF_TILE_MMA)) return 0;
    if (p.b.n != 0 || p.m.n != 1 || p.n.n != 1 || p.k.n != 1) return 0;
    const double *A = (const double *)d.A.data, *B = (const double *)d.B.data;
    const int64_t M = p.m.tot, N = p.n.tot, K = p.k.tot;
    const int64_t sAm = p.m.s[ROLE_A][0], sAk = p.k.s[ROLE_A][0], sBn = p.n.s[ROLE_B][0], sBk = p.k.s[ROLE_B][0];
    const bool    ak = (sAk == 1), bk = (sBk == 1);
    if (!ak && sAm != 1) return 0;
    if (!bk && sBn != 1) return 0;
    const int64_t lda = ak ? sAm : sAk, ldb = bk ? sBn : sBk;
    if (lda <= 0 || ldb <= 0 || (lda & 1) || (ldb & 1) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return 0;
    if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31) || lda >= (1LL << 36) || ldb >= (1LL << 36)) return 0;
    // an extent-1 "other" direction makes ld meaningless but harmless; TMA needs ld*8 to be a multiple of 16 (checked) and >= inner*8
    if ((ak ? K : M) > lda && (ak ? M : K) > 1) return 0;
    if ((bk ? K : N) > ldb && (bk ? N : K) > 1) return 0;
    // split-K and the 80x80 tile live in the generic TMA kernel (fused reduction, no finalize pass)
    const TmaChoice ch = choose_tma_tile((M + 15) / 16, (N + 15) / 16, 1, (K + 15) / 16, d.split_k);
    if (ch.cfg != 0 || ch.splits != 1) return 0;

    CUtensorMap tmA, tmB;
    if (!(ak ? make_operand_map(&tmA, A, K, M, lda, 16, 128) : make_operand_map(&tmA, A, M, K, lda, 16, 16))) return 0;
    if (!(bk ? make_operand_map(&tmB, B, K, N, ldb, 16, 64) : make_operand_map(&tmB, B, N, K, ldb, 16, 16))) return 0;

    TmaGemmArgs a;
    a.M = M; a.N = N; a.K = K;
    a.sCm = p.m.s[ROLE_C][0]; a.sCn = p.n.s[ROLE_C][0];
    a.C = (double *)d.C.data;
    a.alpha = d.alpha[0]; a.beta = beta_eff[0];
    const int64_t tm = (M + 127) / 128, tn = (N + 63) / 64;
    if (tm * tn > 0x7fffffffLL) return 0;
    a.tiles_m = (int)tm; a.tiles_n = (int)tn;
    a.vec_c = (a.sCn == 1 && !(N & 1) && !(a.sCm & 1) && ((uintptr_t)a.C % 16) == 0) ? 1 : 0;
    {
        const int want = option(OPT_TMA_DIST);
        a.dist = (want >= 1 && want < 4) ? want : 3;
        // Rasterisation: groups of `group` tile rows sweep across all tile columns. The A panels of a group stay in L2 for the sweep; the B
        // panels (N x K x 8 bytes: 128 MB for Wabef) do not survive until the next group, so they come from DRAM once per group. 8 rows cost
        // 5 GB of re-reads on Wabef o=20 v=200 (ncu: 30.8 GB moved for 25.9 GB algorithmic); 32 rows keep the live set at ~15 MB and cut
        // the re-reads by four.
        static const int group_env = getenv("EB200_TMA_GROUP") ? atoi(getenv("EB200_TMA_GROUP")) : 0;
        a.group = group_env > 0 ? group_env : 32;
        static const int prefetch_desc = env_int_early("EB200_TMA_PREFETCH_DESC", 0);
        a.prefetch_desc = prefetch_desc;
    }
    d.split_k = 1;
    static const size_t smem_pad = getenv("EB200_TMA_SMEM_PAD") ? (size_t)atoi(getenv("EB200_TMA_SMEM_PAD")) : 0; // debugging: forces 1 CTA/SM
    const size_t        SMEM = 4 * (128 * 16 * 8 + 64 * 16 * 8) + 1024 + 64 + smem_pad;
    {
        int rc_ = k_launch_tma(s, tmA, tmB, a, dim3((unsigned)(tm * tn), 1, 1), SMEM, ak, bk);
        if (rc_) return -rc_;
    }
    d.launches++;
    set_last_kernel("gemm_tma_f64");
    return 1;
}

// ---------------------------------------------------------------- generic TMA path (multi-index groups, batches, split-K, 80x80 tiles)
// One <= 5-dimensional tensor map per operand: dimension 0 is the operand's unit-stride index, the others follow by stride.
// Boxes are 16 x 16 over (inner target index, inner contracted index); see contract_tma.cuh.
struct GenMapDim { int64_t ext, stride; int grp, idx; }; // grp: 0 target (M or N), 1 contracted, 2 batch
bool make_generic_map(CUtensorMap *map, const double *base, const Grp &tgt, const Grp &k, const Grp &b, int role, bool kmajor, int32_t *d_t,
                      int32_t *d_k, int32_t *d_b) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || ((uintptr_t)base % 16)) return false;
    if (tgt.n < 1 || k.n < 1 || tgt.n > TG_MAXI || k.n > TG_MAXI || b.n > TG_MAXI || tgt.n + k.n + b.n > 5) return false;
    std::vector<GenMapDim> dims;
    for (int i = 0; i < tgt.n; i++) dims.push_back({tgt.ext[i], tgt.s[role][i], 0, i});
    for (int i = 0; i < k.n; i++) dims.push_back({k.ext[i], k.s[role][i], 1, i});
    for (int i = 0; i < b.n; i++) dims.push_back({b.ext[i], b.s[role][i], 2, i});
    for (auto &x : dims)
        if (x.stride <= 0 || x.ext < 1 || x.ext >= (1LL << 31) || x.stride >= (1LL << 36)) return false; // broadcast (stride 0) and reversed axes: gather kernel
    const int unit_grp = kmajor ? 1 : 0;
    std::stable_sort(dims.begin(), dims.end(), [&](const GenMapDim &x, const GenMapDim &y) {
        const bool ux = x.grp == unit_grp && x.idx == 0, uy = y.grp == unit_grp && y.idx == 0;
        if (ux != uy) return ux;
        return x.stride < y.stride;
    });
    if (dims[0].stride != 1) return false;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
    int64_t    span = 2;
    for (int i = 0; i < TG_MAXI; i++) d_t[i] = d_k[i] = d_b[i] = -1;
    for (int dd = 0; dd < 5; dd++) {
        if (dd < (int)dims.size()) {
            const GenMapDim &x = dims[dd];
            if (dd > 0 && (x.stride & 1)) return false; // byte strides must be multiples of 16
            gdim[dd] = (cuuint64_t)x.ext;
            if (dd > 0) gstr[dd - 1] = (cuuint64_t)x.stride * 8;
            box[dd] = (x.idx == 0 && x.grp != 2) ? 16 : 1;
            (x.grp == 0 ? d_t : x.grp == 1 ? d_k : d_b)[x.idx] = dd;
            span = std::max(span, x.stride * x.ext);
        } else { // unused dimension: extent 1, coordinate always 0
            gdim[dd] = 1;
            gstr[dd - 1] = (cuuint64_t)((span + 1) & ~1LL) * 8;
            box[dd] = 1;
        }
    }
    for (int dd = 0; dd < 4; dd++)
        if (gstr[dd] >= (1ULL << 40)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void *)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int try_run_tma_generic(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2], bool packed = false) {
    const int use_generic = packed ? 2 : option(OPT_TMA_GENERIC); // packed operands: the caller has applied its own padding rule
    if (!option(OPT_TMA) || !use_generic || (d.force_kernel != 0 && d.force_kernel != EB200_KF_TILE_MMA)) return 0;
    if (p.swapped) return 0;
    const double *A = (const double *)d.A.data, *B = (const double *)d.B.data;
    const bool    ak = p.a_kmajor != 0, bk = p.b_kmajor != 0;
    TmaGenArgs    a;
    memset(&a, 0, sizeof(a));
    CUtensorMap tmA, tmB;
    if (!make_generic_map(&tmA, A, p.m, p.k, p.b, ROLE_A, ak, a.dAm, a.dAk, a.dAb)) return 0;
    if (!make_generic_map(&tmB, B, p.n, p.k, p.b, ROLE_B, bk, a.dBn, a.dBk, a.dBb)) return 0;
    for (int i = 0; i < TG_MAXI; i++) {
        a.em[i] = i < p.m.n ? (int32_t)p.m.ext[i] : 1;
        a.en[i] = i < p.n.n ? (int32_t)p.n.ext[i] : 1;
        a.ek[i] = i < p.k.n ? (int32_t)p.k.ext[i] : 1;
        a.eb[i] = i < p.b.n ? (int32_t)p.b.ext[i] : 1;
        a.sCm[i] = i < p.m.n ? p.m.s[ROLE_C][i] : 0;
        a.sCn[i] = i < p.n.n ? p.n.s[ROLE_C][i] : 0;
        a.sCb[i] = i < p.b.n ? p.b.s[ROLE_C][i] : 0;
    }
    a.mblk0 = (a.em[0] + 15) / 16; a.nblk0 = (a.en[0] + 15) / 16; a.kstep0 = (a.ek[0] + 15) / 16;
    a.MB = (int64_t)a.mblk0 * a.em[1] * a.em[2];
    a.NB = (int64_t)a.nblk0 * a.en[1] * a.en[2];
    const int64_t KT = (int64_t)a.kstep0 * a.ek[1] * a.ek[2];
    if (KT > 0x7fffffffLL) return 0;
    a.KT = (int32_t)KT;
    // padding the inner indices to multiples of 16 must stay cheap, otherwise the gather kernel (no padding) is the better choice
    const double waste = (double)a.MB * 16 / (double)p.m.tot * ((double)a.NB * 16 / (double)p.n.tot) * ((double)KT * 16 / (double)p.k.tot);
    if (waste > 1.12 && use_generic != 2) {
        // a plain product with ragged edges pads only the last block of each direction: always fine
        if (!(p.m.n == 1 && p.n.n == 1 && p.k.n == 1)) return 0;
    }
    const TmaChoice ch = choose_tma_tile(a.MB, a.NB, p.b.tot, KT, d.split_k);
    if (ch.tm * ch.tn > 0x7fffffffLL) return 0;
    a.tiles_m = (int)ch.tm; a.tiles_n = (int)ch.tn;
    a.splits = ch.splits;
    d.split_k = ch.splits;
    a.C = (double *)d.C.data;
    a.alpha = d.alpha[0]; a.beta = beta_eff[0];
    a.vec_c = operand_pairs_ok(p, ROLE_C, p.n, p.m, a.C) ? 1 : 0;
    {
        const int stages = ch.cfg ? TmaTile80x80::STAGES : TmaTile128x64::STAGES, want = option(OPT_TMA_DIST);
        a.dist = (want >= 1 && want < stages) ? want : (stages >= 4 ? stages - 2 : stages - 1);
    }
    const int64_t tiles = ch.tm * ch.tn, per_launch = std::min<int64_t>(65535, p.b.tot);
    const size_t  tile_elems = ch.cfg ? 80 * 80 : 128 * 64;
    void         *ws = nullptr;
    size_t        counter_bytes = 0;
    if (ch.splits > 1) {
        counter_bytes = ((size_t)(tiles * per_launch) * sizeof(unsigned) + 255) & ~(size_t)255;
        int rc = ws_alloc(&ws, counter_bytes + (size_t)(tiles * per_launch) * ch.splits * tile_elems * sizeof(double), s);
        if (rc) return -rc;
        a.counters = (unsigned *)ws;
        a.ws = (double *)((char *)ws + counter_bytes);
    }
    int rc = EB200_OK;
    for (int64_t b0 = 0; b0 < p.b.tot && rc == EB200_OK; b0 += 65535) {
        a.batch0 = b0;
        if (ws) {
            cudaError_t e = cudaMemsetAsync(ws, 0, counter_bytes, s);
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemsetAsync(split counters)"); break; }
        }
        dim3 grid((unsigned)tiles, (unsigned)ch.splits, (unsigned)std::min<int64_t>(65535, p.b.tot - b0));
        rc = k_launch_tma_generic(s, tmA, tmB, a, grid, ch.cfg, ak, bk);
        if (rc == EB200_OK) d.launches++;
    }
    if (ws) {
        cudaError_t fe = cudaFreeAsync(ws, s);
        if (rc == EB200_OK && fe != cudaSuccess) rc = cuda_fail(fe, "cudaFreeAsync");
    }
    if (rc) return -rc;
    set_last_kernel(ch.cfg ? "contract_tma_f64<80x80>" : "contract_tma_f64<128x64>");
    return 1;
}

// FP64 contractions whose own layout has no tensor map (strides that are not multiples of 16 bytes: odd extents) or pads badly in the
// planner's index order: both operands are packed into dense K-major form by one strided copy each (the index with the least padding to
// a multiple of 16 innermost, rows padded to an even length) and the multi-index TMA kernel runs on the packed copies, writing C in
// place through its own strides. The gather kernel it replaces reaches 83 % of the DMMA peak, the TMA kernel 93 %.
int try_run_tma_generic_packed(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    const int use_generic = option(OPT_TMA_GENERIC);
    if (!option(OPT_TMA) || !use_generic || d.force_kernel != 0 || p.swapped) return 0;
    if (p.m.n < 1 || p.n.n < 1 || p.k.n < 1 || p.m.n > TG_MAXI || p.n.n > TG_MAXI || p.b.n > TG_MAXI || p.m.n + p.b.n > 4 || p.n.n + p.b.n > 4) return 0;
    if (p.b.n + p.m.n + p.k.n > EB200_MAX_RANK || p.b.n + p.n.n + p.k.n > EB200_MAX_RANK) return 0;
    for (int i = 0; i < MAXG; i++)
        if ((i < p.m.n && p.m.s[ROLE_A][i] <= 0) || (i < p.n.n && p.n.s[ROLE_B][i] <= 0) || (i < p.k.n && (p.k.s[ROLE_A][i] <= 0 || p.k.s[ROLE_B][i] <= 0)) ||
            (i < p.b.n && (p.b.s[ROLE_A][i] <= 0 || p.b.s[ROLE_B][i] <= 0)))
            return 0;
    const int64_t M = p.m.tot, N = p.n.tot, K = p.k.tot, NBt = p.b.tot, Kp = (K + 1) & ~1LL;
    if (K >= (1LL << 31)) return 0;
    // the two copies cost about 45 (1/M + 1/N) of the product's time; the TMA kernel is 1.24 x the gather kernel on such layouts
    if (use_generic != 2 && (std::min(M, N) < 768 || K < 64)) return 0;
    Plan p2 = p;
    auto reorder = [](Grp &g) { // least padding to a multiple of 16 innermost (all three roles' strides move together)
        int best = 0;
        double bw = 1e30;
        for (int i = 0; i < g.n; i++) {
            const double w = (double)((g.ext[i] + 15) / 16 * 16) / (double)g.ext[i];
            if (w < bw - 1e-12 || (std::fabs(w - bw) <= 1e-12 && g.ext[i] > g.ext[best])) { bw = w; best = i; }
        }
        Grp o = g;
        int c = 0;
        auto put = [&](int i) {
            g.ext[c] = o.ext[i];
            for (int r = 0; r < 3; r++) g.s[r][c] = o.s[r][i];
            c++;
        };
        put(best);
        for (int i = 0; i < o.n; i++)
            if (i != best) put(i);
        grp_finalize(g);
        return bw;
    };
    const double wm = reorder(p2.m), wn = reorder(p2.n);
    const Grp m_src = p2.m, n_src = p2.n; // the ORIGINAL operand strides in the new order (before the packed strides go in)
    // the gather kernel runs odd-stride layouts at 72-75 % of the DMMA peak (no paired loads), the TMA kernel packed operands at 93 %
    if (use_generic != 2 && wm * wn * ((double)((K + 15) / 16 * 16) / (double)K) > 1.25) return 0;
    // packed strides
    int64_t st = Kp;
    for (int i = 0; i < p2.m.n; i++) { p2.m.s[ROLE_A][i] = st; st *= p2.m.ext[i]; }
    for (int i = 0; i < p2.b.n; i++) { p2.b.s[ROLE_A][i] = st; st *= p2.b.ext[i]; }
    const int64_t elemsA = st;
    st = Kp;
    for (int i = 0; i < p2.n.n; i++) { p2.n.s[ROLE_B][i] = st; st *= p2.n.ext[i]; }
    for (int i = 0; i < p2.b.n; i++) { p2.b.s[ROLE_B][i] = st; st *= p2.b.ext[i]; }
    const int64_t elemsB = st;
    memset(&p2.k, 0, sizeof(p2.k));
    p2.k.n = 1; p2.k.tot = K; p2.k.ext[0] = K; p2.k.s[ROLE_A][0] = p2.k.s[ROLE_B][0] = 1;
    grp_finalize(p2.k);
    p2.a_kmajor = p2.b_kmajor = 1;
    void *Ap = nullptr, *Bp = nullptr;
    int   rc;
    scratch_expect((size_t)(elemsA + elemsB) * 8);
    if ((rc = scratch_alloc(&Ap, (size_t)elemsA * 8, s)) || (rc = scratch_alloc(&Bp, (size_t)elemsB * 8, s))) {
        if (Ap) cudaFreeAsync(Ap, s);
        return rc == EB200_ERR_OUT_OF_MEMORY ? 0 : -rc;
    }
    auto pack = [&](const Grp &src, const Grp &dst, int role, const void *from, void *to) -> int {
        int64_t dims[EB200_MAX_RANK], ss[EB200_MAX_RANK], ds[EB200_MAX_RANK];
        int     perm[EB200_MAX_RANK], r = 0;
        int64_t kst = 1;
        for (int i = 0; i < p.k.n; i++, r++) { dims[r] = p.k.ext[i]; ss[r] = p.k.s[role][i]; ds[r] = kst; kst *= p.k.ext[i]; }
        for (int i = 0; i < src.n; i++, r++) { dims[r] = src.ext[i]; ss[r] = src.s[role][i]; ds[r] = dst.s[role][i]; }
        for (int i = 0; i < p.b.n; i++, r++) { dims[r] = p.b.ext[i]; ss[r] = p.b.s[role][i]; ds[r] = p2.b.s[role][i]; }
        for (int i = 0; i < r; i++) perm[i] = i;
        const double one = 1.0, zero = 0.0;
        return eb200_dpermute((void *)s, r, dims, ss, perm, ds, &one, (const double *)from, &zero, (double *)to, 0);
    };
    const int64_t launches0 = launch_counter();
    if ((rc = pack(m_src, p2.m, ROLE_A, d.A.data, Ap)) || (rc = pack(n_src, p2.n, ROLE_B, d.B.data, Bp))) {
        cudaFreeAsync(Ap, s);
        cudaFreeAsync(Bp, s);
        return rc == EB200_ERR_UNSUPPORTED ? 0 : -rc;
    }
    const int packs = (int)(launch_counter() - launches0);
    eb200_contract_desc d2 = d;
    d2.A.data = Ap;
    d2.B.data = Bp;
    d2.launches = 0;
    int r = try_run_tma_generic(s, p2, d2, beta_eff, true);
    cudaFreeAsync(Ap, s);
    cudaFreeAsync(Bp, s);
    if (r != 1) return r; // 0: C untouched, the gather kernel takes over
    d.launches += d2.launches + packs;
    d.split_k = d2.split_k;
    set_last_kernel("contract_tma_f64<packed>");
    return 1;
}

// ---------------------------------------------------------------- tcgen05 3xTF32 path (plain FP32 GEMM, both operands K-major)
int g_use_tcgen05 = -1; // EB200_USE_TCGEN05=0 falls back to the FMA tile kernel
int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
// 3-D map over an FP32 operand: dim0 = the unit-stride direction, dim1 = the other one (ld elements apart), dim2 = batch
bool make_operand_map_f32(CUtensorMap *map, const float *base, int64_t inner, int64_t outer, int64_t ld, int64_t batch, int64_t batch_stride,
                          uint32_t box_inner, uint32_t box_outer, bool mn = false) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if (batch <= 1) batch_stride = std::max<int64_t>(ld * outer, 4); // unused dimension: any legal stride
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)std::max<int64_t>(batch, 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)batch_stride * 4};
    cuuint32_t box[3] = {box_inner, box_outer, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if ((strides[0] & 15) || (strides[1] & 15) || strides[0] >= (1ULL << 40) || strides[1] >= (1ULL << 40)) return false;
    CUtensorMapSwizzle sw = mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B; // see gemm_tf32x3.cuh
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
int try_run_tf32x3(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    if (g_use_tcgen05 < 0) {
        const char *e = getenv("EB200_USE_TCGEN05");
        g_use_tcgen05 = e ? atoi(e) : 1;
    }
    if (!g_use_tcgen05 || d.force_kernel != 0 || d.split_k > 1 || p.swapped) return 0;
    // plain products and products batched over ONE index (strided-batched GEMM, equal-size blocks of the block/tiled algebra)
    if (p.b.n > 1 || p.m.n != 1 || p.n.n != 1 || p.k.n != 1) return 0;
    const float  *A = (const float *)d.A.data, *B = (const float *)d.B.data;
    const int64_t M = p.m.tot, N = p.n.tot, K = p.k.tot, NB = p.b.n ? p.b.tot : 1;
    const int64_t sAm = p.m.s[ROLE_A][0], sAk = p.k.s[ROLE_A][0], sBn = p.n.s[ROLE_B][0], sBk = p.k.s[ROLE_B][0];
    const int64_t sAb = p.b.n ? p.b.s[ROLE_A][0] : 0, sBb = p.b.n ? p.b.s[ROLE_B][0] : 0, sCb = p.b.n ? p.b.s[ROLE_C][0] : 0;
    const bool    a_mn = (sAk != 1), b_mn = (sBk != 1); // MN-major: the target index is the contiguous one
    if (a_mn && sAm != 1) return 0;
    if (b_mn && sBn != 1) return 0;
    const int64_t lda = a_mn ? sAk : sAm, ldb = b_mn ? sBk : sBn;
    if (lda <= 0 || ldb <= 0 || (lda & 3) || (ldb & 3) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return 0;
    if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31) || lda >= (1LL << 36) || ldb >= (1LL << 36)) return 0;
    if ((a_mn ? M : K) > lda || (b_mn ? N : K) > ldb) return 0;
    if (p.b.n && (sAb <= 0 || sBb <= 0 || (sAb & 3) || (sBb & 3) || NB >= (1LL << 31))) return 0; // shared (stride 0) operands: FMA kernel
    if (M * N * NB < 128 * 128 * 4) return 0; // tiny outputs: the FMA kernel (with split-K) is the better fit
    // CTA pairs (cta_group::2, 256 x 128 tile pairs) when the 256-row granularity pads M by at most 5 % and there is work for every pair
    const int64_t tm1 = (M + 127) / 128, tm2 = (M + 255) / 256, tn1 = (N + 127) / 128;
    const int     pair_mode = option(OPT_TCGEN05_PAIR); // 0: never, 1: by the rule, 2: always
    const bool    pair = pair_mode == 2 || (pair_mode == 1 && 2 * tm2 * 20 <= tm1 * 21 && tm2 * tn1 * NB >= sm_count() / 2);
    CUtensorMap tmA, tmB;
    if (!(a_mn ? make_operand_map_f32(&tmA, A, M, K, lda, NB, sAb, 32, 32, true) : make_operand_map_f32(&tmA, A, K, M, lda, NB, sAb, 32, 128))) return 0;
    if (!(b_mn ? make_operand_map_f32(&tmB, B, N, K, ldb, NB, sBb, 32, 32, true) : make_operand_map_f32(&tmB, B, K, N, ldb, NB, sBb, 32, pair ? 64 : 128)))
        return 0;
    Tf32GemmArgs a;
    a.M = M; a.N = N; a.K = K;
    a.sCm = p.m.s[ROLE_C][0]; a.sCn = p.n.s[ROLE_C][0];
    a.C = (float *)d.C.data;
    a.alpha = (float)d.alpha[0]; a.beta = (float)beta_eff[0];
    a.batch = NB; a.sCb = sCb;
    a.c_vec4 = (a.sCn == 1 && (a.sCm & 3) == 0 && (N & 3) == 0 && (sCb & 3) == 0 && ((uintptr_t)d.C.data % 16) == 0) ? 1 : 0;
    const int64_t tm = pair ? tm2 : tm1, tn = tn1; // PAIR: tiles_m counts 256-row tile pairs
    if (tm * tn > 0x7fffffffLL || tm * tn * NB > (1LL << 40)) return 0;
    a.tiles_m = (int)tm; a.tiles_n = (int)tn;
    constexpr size_t SMEM = 3 * 65536 + 4 * (32 * 33 * 4) + 1024 + 256;
    // persistent: one CTA per SM (the kernel needs all of shared memory and all 512 TMEM columns), each walking its share of the tiles
    int rc;
    if (pair) {
        const unsigned grid = 2u * (unsigned)std::min<int64_t>(tm * tn * NB, sm_count() / 2);
        rc = k_launch_tf32x3_pair(s, tmA, tmB, a, dim3(grid, 1, 1), SMEM, a_mn, b_mn);
    } else {
        const unsigned grid = (unsigned)std::min<int64_t>(tm * tn * NB, sm_count());
        rc = k_launch_tf32x3(s, tmA, tmB, a, dim3(grid, 1, 1), SMEM, a_mn, b_mn);
    }
    if (rc) return -rc;
    d.launches++;
    d.split_k = 1;
    d.kernel_used = EB200_KF_TILE_TCGEN05;
    set_last_kernel(pair ? "gemm_tf32x3_tcgen05<pair>" : "gemm_tf32x3_tcgen05");
    return 1;
}

// ---------------------------------------------------------------- complex FP64 on the FP64 tensor pipe (plain products)
// one plain c128 product on gemm_tma_z: 0 = not taken (layout / size), 1 = launched, < 0 = -error
int launch_tma_z(cudaStream_t s, const double2 *A, const double2 *B, double2 *C, int64_t M, int64_t N, int64_t K, int64_t sAm, int64_t sAk, int64_t sBn,
                 int64_t sBk, int64_t sCm, int64_t sCn, const double alpha[2], const double beta[2], int conj_a, int conj_b) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return 0;
    const bool ak = (sAk == 1), bk = (sBk == 1);
    if (!ak && sAm != 1) return 0;
    if (!bk && sBn != 1) return 0;
    const int64_t lda = ak ? sAm : sAk, ldb = bk ? sBn : sBk;
    if (lda <= 0 || ldb <= 0 || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return 0;
    if (M >= (1LL << 30) || N >= (1LL << 30) || K >= (1LL << 30) || lda >= (1LL << 35) || ldb >= (1LL << 35)) return 0;
    if ((ak ? K : M) > lda && (ak ? M : K) > 1) return 0;
    if ((bk ? K : N) > ldb && (bk ? N : K) > 1) return 0;
    const int64_t tm = (M + 127) / 128, tn = (N + 31) / 32;
    if (tm * tn > 0x7fffffffLL) return 0;
    auto make_map = [&](CUtensorMap *map, const double2 *base, int64_t inner, int64_t outer, int64_t ld, uint32_t box_outer) {
        cuuint64_t dims[2] = {(cuuint64_t)inner * 2, (cuuint64_t)outer};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 16};
        cuuint32_t box[2] = {16, box_outer}, es[2] = {1, 1};
        return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    CUtensorMap tmA, tmB;
    if (!(ak ? make_map(&tmA, A, K, M, lda, 128) : make_map(&tmA, A, M, K, lda, 8))) return 0;
    if (!(bk ? make_map(&tmB, B, K, N, ldb, 32) : make_map(&tmB, B, N, K, ldb, 8))) return 0;
    TmaZGemmArgs a;
    a.M = M; a.N = N; a.K = K;
    a.sCm = sCm; a.sCn = sCn;
    a.C = C;
    a.alpha[0] = alpha[0]; a.alpha[1] = alpha[1];
    a.beta[0] = beta[0]; a.beta[1] = beta[1];
    a.tiles_m = (int)tm; a.tiles_n = (int)tn;
    a.conj_a = conj_a; a.conj_b = conj_b;
    const size_t SMEM = 4 * (128 * 8 * 16 + 32 * 8 * 16) + 1024 + 64;
    int rc = k_launch_tma_z(s, tmA, tmB, a, dim3((unsigned)(tm * tn), 1, 1), SMEM, ak, bk);
    return rc ? -rc : 1;
}

int try_run_tma_z(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    if (!option(OPT_TMA) || d.force_kernel != 0 || d.split_k > 1 || p.swapped) return 0;
    if (p.b.n != 0 || p.m.n != 1 || p.n.n != 1 || p.k.n != 1) return 0;
    const int64_t tm = (p.m.tot + 127) / 128, tn = (p.n.tot + 31) / 32;
    // small outputs want split-K (the FMA kernel has it); the tensor-pipe kernel needs enough tiles to fill the machine
    if (tm * tn * 2 <= (int64_t)sm_count() * 2) return 0;
    int r = launch_tma_z(s, (const double2 *)d.A.data, (const double2 *)d.B.data, (double2 *)d.C.data, p.m.tot, p.n.tot, p.k.tot, p.m.s[ROLE_A][0],
                         p.k.s[ROLE_A][0], p.n.s[ROLE_B][0], p.k.s[ROLE_B][0], p.m.s[ROLE_C][0], p.n.s[ROLE_C][0], d.alpha, beta_eff, d.conj_a, d.conj_b);
    if (r != 1) return r;
    d.launches++;
    d.split_k = 1;
    set_last_kernel("gemm_tma_z");
    return 1;
}

// c128 contractions that are not a plain product (multi-index groups, batches, strided layouts): TTGT around gemm_tma_z. A and B are
// gathered into K-major matrices by one strided copy each (alpha and the conjugations folded in), every batch member is one product on
// the FP64 tensor pipe into a dense temporary, and one strided accumulate  C = beta C + T  scatters it. The copies move
// O((M + N) K + M N) elements against 8 M N K flops; the FMA-pipe kernel they replace runs at two thirds of the tensor-pipe rate.
int try_run_z_packed(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    const int mode = option(OPT_TMA_GENERIC); // 2: whenever legal (tests)
    if (!option(OPT_TMA) || !mode || d.force_kernel != 0 || d.split_k > 1 || p.swapped) return 0;
    if (p.m.n < 1 || p.n.n < 1 || p.k.n < 1) return 0;
    if (p.b.n + p.m.n + p.k.n > EB200_MAX_RANK || p.b.n + p.n.n + p.k.n > EB200_MAX_RANK || p.b.n + p.m.n + p.n.n > EB200_MAX_RANK) return 0;
    for (int i = 0; i < MAXG; i++) // broadcast / one-sided sums (stride 0) and negative strides stay on the gather kernel
        if ((i < p.m.n && (p.m.s[ROLE_A][i] <= 0 || p.m.s[ROLE_C][i] <= 0)) || (i < p.n.n && (p.n.s[ROLE_B][i] <= 0 || p.n.s[ROLE_C][i] <= 0)) ||
            (i < p.k.n && (p.k.s[ROLE_A][i] <= 0 || p.k.s[ROLE_B][i] <= 0)) ||
            (i < p.b.n && (p.b.s[ROLE_A][i] <= 0 || p.b.s[ROLE_B][i] <= 0 || p.b.s[ROLE_C][i] <= 0)))
            return 0;
    const int64_t M = p.m.tot, N = p.n.tot, K = p.k.tot, NBt = p.b.tot;
    const int64_t tm = (M + 127) / 128, tn = (N + 31) / 32;
    if (NBt > 256) return 0; // one launch per batch member
    if (mode != 2 && (tm * tn <= sm_count() || K < 32)) return 0;
    void *Ap = nullptr, *Bp = nullptr, *Tp = nullptr;
    int   rc;
    scratch_expect(((size_t)NBt * M * K + (size_t)NBt * N * K + (size_t)NBt * M * N) * 16);
    if ((rc = scratch_alloc(&Ap, (size_t)NBt * M * K * 16, s)) || (rc = scratch_alloc(&Bp, (size_t)NBt * N * K * 16, s)) ||
        (rc = scratch_alloc(&Tp, (size_t)NBt * M * N * 16, s))) {
        if (Ap) cudaFreeAsync(Ap, s);
        if (Bp) cudaFreeAsync(Bp, s);
        return rc == EB200_ERR_OUT_OF_MEMORY ? 0 : -rc; // no room for the packed operands: the gather kernel needs none
    }
    auto release = [&]() {
        cudaFreeAsync(Ap, s);
        cudaFreeAsync(Bp, s);
        cudaFreeAsync(Tp, s);
    };
    // strided copy between a tensor's own layout and the dense [batch][rows][cols] form (inner-first axis lists)
    auto copy = [&](const Grp &rows, const Grp &cols, int role, const void *src, void *dst, bool to_dense, const double *al, const double *be, int conj) -> int {
        int64_t dims[EB200_MAX_RANK], ss[EB200_MAX_RANK], ds[EB200_MAX_RANK];
        int     perm[EB200_MAX_RANK], r = 0;
        int64_t st = 1;
        const Grp *gs[3] = {&cols, &rows, &p.b};
        for (const Grp *g : gs)
            for (int i = 0; i < g->n; i++, r++) {
                dims[r] = g->ext[i];
                (to_dense ? ss : ds)[r] = g->s[role][i];
                (to_dense ? ds : ss)[r] = st;
                st *= g->ext[i];
            }
        for (int i = 0; i < r; i++) perm[i] = i;
        return eb200_zpermute((void *)s, r, dims, ss, perm, ds, al, (const double *)src, be, (double *)dst, conj);
    };
    const double  one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
    const int64_t launches0 = launch_counter();
    if ((rc = copy(p.m, p.k, ROLE_A, d.A.data, Ap, true, d.alpha, zero, d.conj_a)) || (rc = copy(p.n, p.k, ROLE_B, d.B.data, Bp, true, one, zero, d.conj_b))) {
        release();
        return rc == EB200_ERR_UNSUPPORTED ? 0 : -rc; // C is untouched: the gather kernel runs
    }
    for (int64_t b = 0; b < NBt; b++) {
        int r = launch_tma_z(s, (const double2 *)Ap + b * M * K, (const double2 *)Bp + b * N * K, (double2 *)Tp + b * M * N, M, N, K, K, 1, K, 1, N, 1, one, zero, 0, 0);
        if (r != 1) { // nothing of C has been touched yet
            release();
            return r < 0 ? r : 0;
        }
    }
    rc = copy(p.m, p.n, ROLE_C, Tp, d.C.data, false, one, beta_eff, 0);
    release();
    if (rc) return rc == EB200_ERR_UNSUPPORTED ? 0 : -rc; // refused before any launch: C is untouched
    d.launches += (int)(launch_counter() - launches0);
    d.split_k = 1;
    set_last_kernel("gemm_tma_z<packed>");
    return 1;
}

// ---------------------------------------------------------------- generic FP32 contraction on tcgen05 (multi-index groups, batches)
// Same scheme as the FP64 multi-index TMA kernel: one <= 5-D tensor map per operand, 32 x 32 boxes over (the operand's unit-stride
// index, the inner index of the other group); composite M / N / K positions enumerate the INNER index padded to a multiple of 32.
bool make_generic_map_f32(CUtensorMap *map, const float *base, const Grp &tgt, const Grp &k, const Grp &b, int role, bool kmajor, int32_t *d_t,
                          int32_t *d_k, int32_t *d_b) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || ((uintptr_t)base % 16)) return false;
    if (tgt.n < 1 || k.n < 1 || tgt.n > 3 || k.n > 3 || b.n > 3 || tgt.n + k.n + b.n > 5) return false;
    std::vector<GenMapDim> dims;
    for (int i = 0; i < tgt.n; i++) dims.push_back({tgt.ext[i], tgt.s[role][i], 0, i});
    for (int i = 0; i < k.n; i++) dims.push_back({k.ext[i], k.s[role][i], 1, i});
    for (int i = 0; i < b.n; i++) dims.push_back({b.ext[i], b.s[role][i], 2, i});
    for (auto &x : dims)
        if (x.stride <= 0 || x.ext < 1 || x.ext >= (1LL << 31) || x.stride >= (1LL << 36)) return false;
    const int unit_grp = kmajor ? 1 : 0;
    std::stable_sort(dims.begin(), dims.end(), [&](const GenMapDim &x, const GenMapDim &y) {
        const bool ux = x.grp == unit_grp && x.idx == 0, uy = y.grp == unit_grp && y.idx == 0;
        if (ux != uy) return ux;
        return x.stride < y.stride;
    });
    if (dims[0].stride != 1) return false;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
    int64_t    span = 4;
    for (int i = 0; i < 3; i++) d_t[i] = d_k[i] = d_b[i] = -1;
    for (int dd = 0; dd < 5; dd++) {
        if (dd < (int)dims.size()) {
            const GenMapDim &x = dims[dd];
            if (dd > 0 && (x.stride & 3)) return false; // byte strides must be multiples of 16
            gdim[dd] = (cuuint64_t)x.ext;
            if (dd > 0) gstr[dd - 1] = (cuuint64_t)x.stride * 4;
            box[dd] = (x.idx == 0 && x.grp != 2) ? 32 : 1;
            (x.grp == 0 ? d_t : x.grp == 1 ? d_k : d_b)[x.idx] = dd;
            span = std::max(span, x.stride * x.ext);
        } else {
            gdim[dd] = 1;
            gstr[dd - 1] = (cuuint64_t)((span + 3) & ~3LL) * 4;
            box[dd] = 1;
        }
    }
    for (int dd = 0; dd < 4; dd++)
        if (gstr[dd] >= (1ULL << 40)) return false;
    // K-major tiles: 128-byte rows along k, plain 128-byte swizzle; MN-major tiles: rows along m/n need the 32-byte-atom swizzle
    const CUtensorMapSwizzle sw = kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void *)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- C layouts the tile epilogue cannot write in memory order. The epilogue's lanes run along the INNER index of the M (or N) group; when
// neither is unit-stride in C (Wmbej: W[m,b,e,j] has j innermost, an OUTER index of the M group), every 4-byte access of a warp falls into its own
// 32-byte sector. With beta = 0 the L2 merges the partial stores (188 TFLOP/s on Wmbej FP32 o=40 v=160); with beta != 0 the old-C loads cost 8 x
// their bytes and throttle the kernel (129). Then the tile kernel writes a dense temporary in GEMM order (lanes along m, unit stride) with
// beta = 0, and one strided accumulate  C = beta C + T  (the permute kernels: both sides coalesced through shared-memory tiles) finishes: 192.
struct DenseOut {
    void   *T = nullptr;
    int     n = 0;
    int64_t ext[EB200_MAX_RANK], ts[EB200_MAX_RANK], cs[EB200_MAX_RANK];
};
// fills strides of the dense temporary into gCm / gCn / gCb (in FLOATS, w floats per element) and records the index list for the accumulate
int dense_out_begin(cudaStream_t s, DenseOut &o, int w, const int32_t em[3], int nm, const int32_t en[3], int nn, const int32_t eb[3], int nb, int64_t gCm[3],
                    int64_t gCn[3], int64_t gCb[3]) {
    int64_t st = 1;
    o.n = 0;
    auto add = [&](const int32_t *e, int cnt, int64_t *g) {
        for (int i = 0; i < cnt; i++) {
            o.ext[o.n] = e[i];
            o.cs[o.n] = g[i] / w; // the caller's C strides, back in elements
            o.ts[o.n] = st;
            g[i] = st * w;
            st *= e[i];
            o.n++;
        }
    };
    add(em, nm, gCm);
    add(en, nn, gCn);
    add(eb, nb, gCb);
    int rc = scratch_alloc(&o.T, (size_t)st * 4 * w, s);
    return rc;
}
int dense_out_finish(cudaStream_t s, DenseOut &o, bool cplx, void *C, const double beta[2]) {
    int perm[EB200_MAX_RANK];
    for (int i = 0; i < o.n; i++) perm[i] = i;
    const float one[2] = {1.f, 0.f}, be[2] = {(float)beta[0], (float)beta[1]};
    int rc = cplx ? eb200_cpermute((void *)s, o.n, o.ext, o.ts, perm, o.cs, one, (const float *)o.T, be, (float *)C, 0)
                  : eb200_spermute((void *)s, o.n, o.ext, o.ts, perm, o.cs, one, (const float *)o.T, be, (float *)C, 0);
    cudaFreeAsync(o.T, s);
    o.T = nullptr;
    return rc;
}

int try_run_tf32x3_generic(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    if (g_use_tcgen05 < 0) {
        const char *e = getenv("EB200_USE_TCGEN05");
        g_use_tcgen05 = e ? atoi(e) : 1;
    }
    const int mode = option(OPT_TCGEN05_GENERIC); // 1: by the rules below; 0: never; 2: whenever a tensor map exists (tests); 3: packed operands preferred
    if (!g_use_tcgen05 || !mode || mode == 3 || d.force_kernel != 0 || d.split_k > 1 || p.swapped) return 0;
    const float *A = (const float *)d.A.data, *B = (const float *)d.B.data;
    const bool   ak = p.a_kmajor != 0, bk = p.b_kmajor != 0;
    Tf32GemmArgs a;
    memset(&a, 0, sizeof(a));
    CUtensorMap tmA, tmB;
    if (!make_generic_map_f32(&tmA, A, p.m, p.k, p.b, ROLE_A, ak, a.dAm, a.dAk, a.dAb)) return 0;
    if (!make_generic_map_f32(&tmB, B, p.n, p.k, p.b, ROLE_B, bk, a.dBn, a.dBk, a.dBb)) return 0;
    for (int i = 0; i < 3; i++) {
        a.em[i] = i < p.m.n ? (int32_t)p.m.ext[i] : 1;
        a.en[i] = i < p.n.n ? (int32_t)p.n.ext[i] : 1;
        a.ek[i] = i < p.k.n ? (int32_t)p.k.ext[i] : 1;
        a.eb[i] = i < p.b.n ? (int32_t)p.b.ext[i] : 1;
        a.gCm[i] = i < p.m.n ? p.m.s[ROLE_C][i] : 0;
        a.gCn[i] = i < p.n.n ? p.n.s[ROLE_C][i] : 0;
        a.gCb[i] = i < p.b.n ? p.b.s[ROLE_C][i] : 0;
    }
    a.mblk0 = (a.em[0] + 31) / 32; a.nblk0 = (a.en[0] + 31) / 32; a.kstep0 = (a.ek[0] + 31) / 32;
    a.MB = (int64_t)a.mblk0 * a.em[1] * a.em[2];
    a.NB = (int64_t)a.nblk0 * a.en[1] * a.en[2];
    const int64_t KT = (int64_t)a.kstep0 * a.ek[1] * a.ek[2];
    if (KT > 0x7fffffffLL || p.b.tot >= (1LL << 31)) return 0;
    a.KT = (int32_t)KT;
    // padding the inner indices to multiples of 32 costs tensor-pipe work; the FMA-pipe kernel (no padding) is ~5 x slower, so a factor
    // of 2.5 is still a clear win
    const double waste = (double)a.MB * 32 / (double)p.m.tot * ((double)a.NB * 32 / (double)p.n.tot) * ((double)KT * 32 / (double)p.k.tot);
    if (waste > 2.5 && mode != 2) return 0;
    int64_t       tm = (a.MB + 3) / 4;
    const int64_t tn = (a.NB + 3) / 4;
    if (tm * tn > 0x7fffffffLL || tm * tn * p.b.tot > (1LL << 40)) return 0;
    if (tm * tn * p.b.tot * 4 < sm_count() && mode != 2) return 0; // too few tiles to fill the machine: the FMA kernel has split-K
    // CTA pairs (256-row tile pairs) by the rule of the plain kernel
    const int64_t tm2 = (a.MB + 7) / 8;
    const int     pair_mode = option(OPT_TCGEN05_PAIR);
    const bool    pair = pair_mode == 2 || (pair_mode == 1 && 2 * tm2 * 20 <= tm * 21 && tm2 * tn * p.b.tot >= sm_count() / 2);
    if (pair) tm = tm2;
    a.tiles_m = (int)tm; a.tiles_n = (int)tn;
    a.batch = p.b.tot;
    a.M = p.m.tot; a.N = p.n.tot; a.K = p.k.tot;
    a.C = (float *)d.C.data;
    a.alpha = (float)d.alpha[0]; a.beta = (float)beta_eff[0];
    a.c_n_unit = (p.n.n > 0 && p.n.s[ROLE_C][0] == 1) ? 1 : 0;
    DenseOut dense;
    if (beta_eff[0] != 0.0 && !a.c_n_unit && p.m.s[ROLE_C][0] != 1 && (mode == 2 || (double)p.m.tot * (double)p.n.tot * (double)p.b.tot >= (double)(1 << 22))) {
        int rc = dense_out_begin(s, dense, 1, a.em, p.m.n, a.en, p.n.n, a.eb, p.b.n, a.gCm, a.gCn, a.gCb);
        if (rc == EB200_OK) {
            a.C = (float *)dense.T;
            a.beta = 0.f;
        } else if (rc != EB200_ERR_OUT_OF_MEMORY) return -rc;
        else // no room: write C directly (slower, still correct)
            for (int i = 0; i < 3; i++) {
                a.gCm[i] = i < p.m.n ? p.m.s[ROLE_C][i] : 0;
                a.gCn[i] = i < p.n.n ? p.n.s[ROLE_C][i] : 0;
                a.gCb[i] = i < p.b.n ? p.b.s[ROLE_C][i] : 0;
            }
    }
    constexpr size_t SMEM = 3 * 65536 + 4 * (32 * 33 * 4) + 1024 + 256 + 2048;
    const unsigned   grid = pair ? 2u * (unsigned)std::min<int64_t>(tm * tn * p.b.tot, sm_count() / 2) : (unsigned)std::min<int64_t>(tm * tn * p.b.tot, sm_count());
    int rc = k_launch_tf32x3_generic(s, tmA, tmB, a, dim3(grid, 1, 1), SMEM, !ak, !bk, pair);
    if (rc) {
        if (dense.T) cudaFreeAsync(dense.T, s);
        return -rc;
    }
    if (dense.T) {
        const int64_t l0 = launch_counter();
        rc = dense_out_finish(s, dense, false, d.C.data, beta_eff);
        if (rc) return -rc;
        d.launches += (int)(launch_counter() - l0);
    }
    d.launches++;
    d.split_k = 1;
    d.kernel_used = EB200_KF_TILE_TCGEN05;
    set_last_kernel(pair ? "gemm_tf32x3_tcgen05<generic,pair>" : "gemm_tf32x3_tcgen05<generic>");
    return 1;
}

// ---------------------------------------------------------------- c64 contractions on tcgen05 (through the real 3xTF32 kernel)
// A complex product is a real product of twice the size: with A' = A viewed as real rows [re, im interleaved along k] and
//     B'[re-plane][n] = (Bre, -Bim ...),  B'[im-plane][n] = (Bim, Bre ...)
// the real GEMM  A' (M x 2K) . B'^T (2K x 2N)  yields the real parts in the first plane's columns and the imaginary parts in the second
// plane's — 8 M N K real flops, exactly the complex product's count. Both operands are first PACKED (one strided-copy launch each: the
// index permutation, conjugation and alpha are absorbed there), so every tile-family pattern becomes one batched K-major product whose
// multi-index C is written by the generic epilogue as (re, im) pairs. Packing moves O((M + N) K) bytes against O(M N K) flops.
// CPLX = false is the FP32 variant of the same scheme (no planes, no expansion): it serves real contractions whose own layout has no
// tensor map (strides that are not multiples of 16 bytes, no unit-stride index) and would otherwise run on the FMA pipe.
template <bool CPLX> int try_run_packed_tcgen05(cudaStream_t s, const Plan &p, eb200_contract_desc &d, double beta_eff[2]) {
    constexpr int64_t W = CPLX ? 2 : 1; // floats per element
    if (g_use_tcgen05 < 0) g_use_tcgen05 = env_int("EB200_USE_TCGEN05", 1);
    const int mode = option(OPT_TCGEN05_GENERIC);
    if (!g_use_tcgen05 || !mode || d.force_kernel != 0 || d.split_k > 1 || p.swapped) return 0;
    if (p.m.n < 1 || p.n.n < 1 || p.k.n < 1 || p.m.n > 3 || p.n.n > (CPLX ? 2 : 3) || p.b.n > 3) return 0;
    if (p.b.n + p.m.n + p.k.n > EB200_MAX_RANK || p.b.n + p.n.n + p.k.n > EB200_MAX_RANK) return 0;
    if (!encode_tiled_fn()) return 0;
    for (int i = 0; i < MAXG; i++) // broadcast / one-sided sums (stride 0) and negative strides stay on the gather kernel
        if ((i < p.m.n && p.m.s[ROLE_A][i] <= 0) || (i < p.n.n && p.n.s[ROLE_B][i] <= 0) || (i < p.k.n && (p.k.s[ROLE_A][i] <= 0 || p.k.s[ROLE_B][i] <= 0)) ||
            (i < p.b.n && (p.b.s[ROLE_A][i] <= 0 || p.b.s[ROLE_B][i] <= 0)))
            return 0;
    const int64_t M = p.m.tot, N = p.n.tot, K = p.k.tot, NBt = p.b.tot, Kp = CPLX ? ((K + 1) & ~1LL) : ((K + 3) & ~3LL); // packed rows: multiples of 16 bytes
    if (W * K >= (1LL << 31) || NBt >= (1LL << 31) || M >= (1LL << 40) || N >= (1LL << 40)) return 0;
    // enumeration order of the M and N groups: the index that pads least to a multiple of 32 goes innermost
    auto order = [](const Grp &g, int *ord) {
        int best = 0;
        double bw = 1e30;
        for (int i = 0; i < g.n; i++) {
            const double w = (double)((g.ext[i] + 31) / 32 * 32) / (double)g.ext[i];
            if (w < bw - 1e-12 || (std::fabs(w - bw) <= 1e-12 && g.ext[i] > g.ext[best])) { bw = w; best = i; }
        }
        int c = 0;
        ord[c++] = best;
        for (int i = 0; i < g.n; i++)
            if (i != best) ord[c++] = i;
    };
    int om[3] = {0, 0, 0}, on[3] = {0, 0, 0};
    order(p.m, om);
    order(p.n, on);
    Tf32GemmArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 3; i++) {
        a.em[i] = i < p.m.n ? (int32_t)p.m.ext[om[i]] : 1;
        a.en[i] = i < p.n.n ? (int32_t)p.n.ext[on[i]] : 1;
        a.ek[i] = 1;
        a.eb[i] = i < p.b.n ? (int32_t)p.b.ext[i] : 1;
        a.gCm[i] = i < p.m.n ? W * p.m.s[ROLE_C][om[i]] : 0; // C strides in floats
        a.gCn[i] = i < p.n.n ? W * p.n.s[ROLE_C][on[i]] : 0;
        a.gCb[i] = i < p.b.n ? W * p.b.s[ROLE_C][i] : 0;
        if ((i < p.m.n && p.m.ext[om[i]] >= (1LL << 31)) || (i < p.n.n && p.n.ext[on[i]] >= (1LL << 31))) return 0;
    }
    a.ek[0] = (int32_t)(W * K);
    a.mblk0 = (a.em[0] + 31) / 32; a.nblk0 = (a.en[0] + 31) / 32;
    a.kstep0 = (int32_t)((W * K + 31) / 32);
    a.KT = a.kstep0;
    a.MB = (int64_t)a.mblk0 * a.em[1] * a.em[2];
    a.NB = (int64_t)a.nblk0 * a.en[1] * a.en[2];
    const int64_t NT = CPLX ? 64 : 128, tm = (a.MB + 3) / 4, tn = CPLX ? (a.NB + 1) / 2 : (a.NB + 3) / 4; // 128 rows x 64 complex (128 real) columns per tile
    if (tm * tn > 0x7fffffffLL || tm * tn * NBt > (1LL << 40)) return 0;
    if (mode != 2) {
        const double waste = (double)(tm * 128) / (double)M * ((double)(tn * NT) / (double)N);
        if (waste > 2.0) return 0;
        if (tm * tn * NBt * (CPLX ? 2 : 4) < sm_count()) return 0;      // too few tiles: the FMA-pipe kernel has split-K
        if (std::min(M, N) < 128 || K < 16) return 0;       // packing would not pay
    }
    // ---- pack: A' [b][m composite][Kp] complex (alpha and conj folded in), Bc [b][n composite][K] complex, B' two planes of rows
    void  *Ap = nullptr, *Bc = nullptr, *Bp = nullptr;
    int    rc = EB200_OK;
    auto   pack = [&](const Grp &tgt, const int *ord, int role, const void *src, void *dst, int64_t row_len, const float *al, int conj) -> int {
        int64_t dims[EB200_MAX_RANK], ss[EB200_MAX_RANK], ds[EB200_MAX_RANK];
        int     perm[EB200_MAX_RANK], r = 0;
        int64_t st = 1;
        for (int i = 0; i < p.k.n; i++, r++) { dims[r] = p.k.ext[i]; ss[r] = p.k.s[role][i]; ds[r] = st; st *= p.k.ext[i]; }
        st = row_len;
        for (int i = 0; i < tgt.n; i++, r++) { dims[r] = tgt.ext[ord[i]]; ss[r] = tgt.s[role][ord[i]]; ds[r] = st; st *= tgt.ext[ord[i]]; }
        for (int i = 0; i < p.b.n; i++, r++) { dims[r] = p.b.ext[i]; ss[r] = p.b.s[role][i]; ds[r] = st; st *= p.b.ext[i]; }
        for (int i = 0; i < r; i++) perm[i] = i;
        const float zero[2] = {0.f, 0.f};
        if constexpr (CPLX) return eb200_cpermute((void *)s, r, dims, ss, perm, ds, al, (const float *)src, zero, (float *)dst, conj);
        else return eb200_spermute((void *)s, r, dims, ss, perm, ds, al, (const float *)src, zero, (float *)dst, 0);
    };
    const float alpha_c[2] = {(float)d.alpha[0], (float)d.alpha[1]}, one_c[2] = {1.f, 0.f};
    const size_t bytesA = (size_t)NBt * M * Kp * 4 * W, bytesBc = CPLX ? (size_t)NBt * N * K * 8 : 0, bytesBp = (size_t)NBt * W * N * Kp * 4 * W;
    scratch_expect(bytesA + bytesBc + bytesBp + (size_t)NBt * M * N * 4 * W); // + the dense temporary some C layouts need
    if ((rc = scratch_alloc(&Ap, bytesA, s)) || (CPLX && (rc = scratch_alloc(&Bc, bytesBc, s))) || (rc = scratch_alloc(&Bp, bytesBp, s))) {
        if (Ap) cudaFreeAsync(Ap, s);
        if (Bc) cudaFreeAsync(Bc, s);
        return rc == EB200_ERR_OUT_OF_MEMORY ? 0 : -rc; // no room for the packed operands: the gather kernel needs none
    }
    auto release = [&]() {
        cudaFreeAsync(Ap, s);
        if (Bc) cudaFreeAsync(Bc, s);
        cudaFreeAsync(Bp, s);
    };
    int64_t launches0 = launch_counter();
    if ((rc = pack(p.m, om, ROLE_A, d.A.data, Ap, Kp, alpha_c, d.conj_a)) ||
        (rc = CPLX ? pack(p.n, on, ROLE_B, d.B.data, Bc, K, one_c, d.conj_b) : pack(p.n, on, ROLE_B, d.B.data, Bp, Kp, one_c, 0)) ||
        (CPLX && (rc = k_launch_c64_expand(s, Bc, Bp, NBt * N, N, K, Kp)))) {
        release();
        return rc == EB200_ERR_UNSUPPORTED ? 0 : -rc; // a layout the strided copy does not take: C is untouched, the gather kernel runs
    }
    // ---- tensor maps over the packed operands (real view: rows of 2 Kp floats, 2 K of them valid — the rest is out of bounds = zero)
    Grp gk, gb, gm, gn;
    memset(&gk, 0, sizeof(gk)); memset(&gb, 0, sizeof(gb)); memset(&gm, 0, sizeof(gm)); memset(&gn, 0, sizeof(gn));
    gk.n = 1; gk.ext[0] = W * K; gk.s[ROLE_A][0] = gk.s[ROLE_B][0] = 1;
    int64_t st = W * Kp;
    gm.n = p.m.n;
    for (int i = 0; i < p.m.n; i++) { gm.ext[i] = a.em[i]; gm.s[ROLE_A][i] = st; st *= a.em[i]; }
    if (NBt > 1) { gb.n = 1; gb.ext[0] = NBt; gb.s[ROLE_A][0] = st; }
    st = W * Kp;
    gn.n = p.n.n + (CPLX ? 1 : 0);
    for (int i = 0; i < p.n.n; i++) { gn.ext[i] = a.en[i]; gn.s[ROLE_B][i] = st; st *= a.en[i]; }
    if (CPLX) { gn.ext[p.n.n] = 2; gn.s[ROLE_B][p.n.n] = st; st *= 2; } // the (real | imaginary)-output planes
    if (NBt > 1) gb.s[ROLE_B][0] = st;
    CUtensorMap tmA, tmB;
    int32_t     dBn[3];
    if (!make_generic_map_f32(&tmA, (const float *)Ap, gm, gk, gb, ROLE_A, true, a.dAm, a.dAk, a.dAb) ||
        !make_generic_map_f32(&tmB, (const float *)Bp, gn, gk, gb, ROLE_B, true, dBn, a.dBk, a.dBb)) {
        release(); // nothing of C has been touched yet: the FMA-pipe kernel takes over
        return 0;
    }
    for (int i = 0; i < 3; i++) a.dBn[i] = i < p.n.n ? dBn[i] : -1;
    a.dBc = CPLX ? dBn[p.n.n] : -1;
    a.cplx = CPLX ? 1 : 0;
    a.b_flat = 1;
    // C layouts the epilogue cannot write in memory order, beta != 0: dense temporary + one strided accumulate (see DenseOut); this also
    // takes a complex beta as it is
    a.c_n_unit = (p.n.s[ROLE_C][on[0]] == 1 && p.m.s[ROLE_C][om[0]] != 1) ? 1 : 0;
    DenseOut dense;
    if ((beta_eff[0] != 0.0 || beta_eff[1] != 0.0) && !a.c_n_unit && p.m.s[ROLE_C][om[0]] != 1 && (mode == 2 || (double)M * (double)N * (double)NBt >= (double)(1 << 22))) {
        rc = dense_out_begin(s, dense, (int)W, a.em, p.m.n, a.en, p.n.n, a.eb, p.b.n, a.gCm, a.gCn, a.gCb);
        if (rc != EB200_OK && rc != EB200_ERR_OUT_OF_MEMORY) { release(); return -rc; }
        if (rc == EB200_ERR_OUT_OF_MEMORY) { // no room: write C directly
            dense.T = nullptr;
            for (int i = 0; i < 3; i++) {
                a.gCm[i] = i < p.m.n ? W * p.m.s[ROLE_C][om[i]] : 0;
                a.gCn[i] = i < p.n.n ? W * p.n.s[ROLE_C][on[i]] : 0;
                a.gCb[i] = i < p.b.n ? W * p.b.s[ROLE_C][i] : 0;
            }
        }
    }
    // the epilogue multiplies by a REAL beta per component; a beta with an imaginary part scales C beforehand
    if (!dense.T && CPLX && beta_eff[1] != 0.0) {
        rc = launch_scale<float2>(s, p.t, d.C.data, beta_eff, d);
        if (rc) { release(); return -rc; }
        d.launches--; // counted below with the other launches of this call
        beta_eff[0] = 1.0; beta_eff[1] = 0.0;
    }
    const int64_t tm2 = (a.MB + 7) / 8;
    const int     pair_mode = option(OPT_TCGEN05_PAIR);
    const bool    pair = pair_mode == 2 || (pair_mode == 1 && 2 * tm2 * 20 <= tm * 21 && tm2 * tn * NBt >= sm_count() / 2);
    a.tiles_m = (int)(pair ? tm2 : tm); a.tiles_n = (int)tn;
    a.batch = NBt;
    a.M = M; a.N = W * N; a.K = W * K;
    a.C = dense.T ? (float *)dense.T : (float *)d.C.data;
    a.alpha = 1.f; a.beta = dense.T ? 0.f : (float)beta_eff[0];
    if (dense.T) a.c_n_unit = 0; // the temporary is unit-stride along the inner M index
    constexpr size_t SMEM = 3 * 65536 + 4 * (32 * 33 * 4) + 1024 + 256 + 2048;
    const unsigned   grid = pair ? 2u * (unsigned)std::min<int64_t>(tm2 * tn * NBt, sm_count() / 2) : (unsigned)std::min<int64_t>(tm * tn * NBt, sm_count());
    rc = k_launch_tf32x3_generic(s, tmA, tmB, a, dim3(grid, 1, 1), SMEM, false, false, pair);
    release();
    if (rc) {
        if (dense.T) cudaFreeAsync(dense.T, s);
        return -rc;
    }
    if (dense.T && (rc = dense_out_finish(s, dense, CPLX, d.C.data, beta_eff))) return -rc;
    d.launches += (int)(launch_counter() - launches0);
    d.split_k = 1;
    d.kernel_used = EB200_KF_TILE_TCGEN05;
    set_last_kernel(CPLX ? "gemm_tf32x3_tcgen05<c64>" : "gemm_tf32x3_tcgen05<packed>");
    return 1;
}

template <typename T> int run_tile(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    TileArgs args;
    args.p = p;
    args.A = p.swapped ? d.B.data : d.A.data;
    args.B = p.swapped ? d.A.data : d.B.data;
    args.C = d.C.data;
    args.ws = nullptr;
    args.alpha[0] = d.alpha[0]; args.alpha[1] = d.alpha[1];
    args.beta[0] = beta_eff[0]; args.beta[1] = beta_eff[1];
    args.conj_a = d.conj_a; args.conj_b = d.conj_b;
    args.vec_ab = 0; args.vec_c = 0;

    const bool mma = (p.family == EB200_KF_TILE_MMA);
    if constexpr (std::is_same_v<T, double>) {
        if (mma) {
            int r = try_run_tma(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
            r = try_run_tma_generic(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
            r = try_run_tma_generic_packed(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
        }
    }
    if constexpr (std::is_same_v<T, double2>) {
        if (!mma) {
            int r = try_run_tma_z(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
            r = try_run_z_packed(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
        }
    }
    if constexpr (std::is_same_v<T, float2>) {
        if (!mma) {
            double be[2] = {beta_eff[0], beta_eff[1]};
            int    r = try_run_packed_tcgen05<true>(s, p, d, be);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
        }
    }
    if constexpr (std::is_same_v<T, float>) {
        if (!mma) {
            int r = try_run_tf32x3(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
            r = try_run_tf32x3_generic(s, p, d, beta_eff);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
            double be[2] = {beta_eff[0], 0.0};
            r = try_run_packed_tcgen05<false>(s, p, d, be);
            if (r == 1) return EB200_OK;
            if (r < 0) return -r;
        }
    }
    int bm, bn, bk, ctas_per_sm;
    std::call_once(g_cfg_once, [] {
        const char *e = getenv("EB200_MMA_CONFIG");
        if (e) g_mma_config = atoi(e);
    });
    int cfg = 0;
    if (mma) {
        cfg = 2; bm = 128; bn = 64; bk = 16; ctas_per_sm = 2;
        args.vec_ab = operand_pairs_ok(p, ROLE_A, p.a_kmajor ? p.k : p.m, p.a_kmajor ? p.m : p.k, args.A) &&
                      operand_pairs_ok(p, ROLE_B, p.b_kmajor ? p.k : p.n, p.b_kmajor ? p.n : p.k, args.B);
        args.vec_c = operand_pairs_ok(p, ROLE_C, p.n, p.m, args.C);
    } else { bm = SimtCfg<T>::BM; bn = SimtCfg<T>::BN; bk = SimtCfg<T>::BK; ctas_per_sm = 2; }

    const int64_t tm = (p.m.tot + bm - 1) / bm, tn = (p.n.tot + bn - 1) / bn;
    if (tm * tn > 0x7fffffffLL)
        return fail(EB200_ERR_UNSUPPORTED, "contract: tile grid too large (%lld x %lld tiles)", (long long)tm, (long long)tn);
    args.tiles_m = (int)tm;
    args.tiles_n = (int)tn;

    // split-K when the output alone cannot fill the machine
    const int64_t tiles = tm * tn * p.b.tot;
    const int64_t KT = (p.k.tot + bk - 1) / bk;
    int           splits = 1;
    if (d.split_k > 0) splits = d.split_k;
    else {
        const int64_t target = (int64_t)sm_count() * ctas_per_sm;
        if (tiles * 2 <= target) splits = (int)std::min<int64_t>(target / tiles, std::max<int64_t>(1, KT / 8));
    }
    splits = (int)std::max<int64_t>(1, std::min<int64_t>(splits, KT));
    if (splits > 65535) splits = 65535;
    args.splits = splits;
    d.split_k = splits;

    void *ws = nullptr;
    if (splits > 1) {
        size_t bytes = (size_t)splits * p.b.tot * p.m.tot * p.n.tot * sizeof(T);
        int    rc = ws_alloc(&ws, bytes, s);
        if (rc) return rc;
        args.ws = ws;
    }
    int rc = EB200_OK;
    for (int64_t b0 = 0; b0 < p.b.tot && rc == EB200_OK; b0 += 65535) { // grid.z holds at most 65535 batches per launch
        dim3 grid((unsigned)(tm * tn), (unsigned)splits, (unsigned)std::min<int64_t>(65535, p.b.tot - b0));
        args.batch0 = b0;
        if (mma) {
            if constexpr (std::is_same_v<T, double>) rc = k_launch_mma(s, args, grid);
            else rc = fail(EB200_ERR_UNSUPPORTED, "tile_mma is FP64 only");
        } else {
            rc = k_launch_simt(Scalar<T>::dtype, s, args, grid);
        }
        if (rc == EB200_OK) d.launches++;
    }
    if (rc) return rc;
    set_last_kernel(mma ? "tile_mma_f64" : "tile_simt");
    if (splits > 1) {
        rc = launch_finalize<T>(s, p.b, p.m, p.n, ws, d.C.data, d.alpha, beta_eff, splits, d);
        if (rc) return rc;
        EB_CUDA(cudaFreeAsync(ws, s));
    }
    return EB200_OK;
}

// ---------------------------------------------------------------- streaming reductions
template <typename T> int run_reduce(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    ReduceArgs a;
    a.rows = p.rows; a.n = p.n; a.k = p.k;
    a.A = p.swapped ? d.B.data : d.A.data;
    a.B = p.swapped ? d.A.data : d.B.data;
    a.C = d.C.data;
    a.ws = nullptr;
    a.alpha[0] = d.alpha[0]; a.alpha[1] = d.alpha[1];
    a.beta[0] = beta_eff[0]; a.beta[1] = beta_eff[1];
    a.conj_a = p.swapped ? d.conj_b : d.conj_a;
    a.conj_b = p.swapped ? d.conj_a : d.conj_b;

    const bool    row = (p.family != EB200_KF_COLREDUCE);
    const int64_t R = p.rows.tot, K = p.k.tot;
    // whole-tensor inner product of two unit-stride operands: dedicated streaming kernel (8 pairs per thread in flight)
    if (row && R == 1 && p.n.tot == 1 && p.k.n == 1 && p.k.s[ROLE_A][0] == 1 && p.k.s[ROLE_B][0] == 1 && K >= 65536 && d.split_k <= 0) {
        const int splits = (int)std::min<int64_t>((int64_t)sm_count() * 8, K / 16384);
        a.splits = splits;
        d.split_k = splits;
        void *ws = nullptr;
        int   rc = ws_alloc(&ws, (size_t)splits * sizeof(T), s);
        if (rc) return rc;
        a.ws = ws;
        rc = k_launch_dot(Scalar<T>::dtype, s, a, splits);
        if (rc) return rc;
        d.launches++;
        Grp none;
        memset(&none, 0, sizeof(none));
        none.tot = 1;
        rc = launch_finalize<T>(s, none, p.rows, p.n, ws, d.C.data, d.alpha, beta_eff, splits, d);
        if (rc) return rc;
        EB_CUDA(cudaFreeAsync(ws, s));
        return EB200_OK;
    }
    const int64_t rows_per_block = row ? reduce_rows_per_block(p.n.tot, p.k.tot) : 256 * 2;
    const int64_t row_blocks = (R + rows_per_block - 1) / rows_per_block;
    if (row_blocks > 0x7fffffffLL) return fail(EB200_ERR_UNSUPPORTED, "contract: too many rows for the reduction grid");
    int splits = 1;
    if (d.split_k > 0) splits = d.split_k;
    else {
        const int64_t target = (int64_t)sm_count() * 8;
        const int64_t min_k_per_split = row ? 8192 : 64;
        if (row_blocks < target) splits = (int)std::min<int64_t>((target + row_blocks - 1) / row_blocks, std::max<int64_t>(1, K / min_k_per_split));
    }
    splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(splits, K), 65535));
    a.splits = splits;
    d.split_k = splits;
    void *ws = nullptr;
    if (splits > 1) {
        int rc = ws_alloc(&ws, (size_t)splits * R * p.n.tot * sizeof(T), s);
        if (rc) return rc;
        a.ws = ws;
    }
    dim3 grid((unsigned)row_blocks, (unsigned)splits, 1);
    int  rc = k_launch_reduce(Scalar<T>::dtype, s, row, a, grid);
    if (rc) return rc;
    d.launches++;
    if (splits > 1) {
        Grp none;
        memset(&none, 0, sizeof(none));
        none.tot = 1;
        rc = launch_finalize<T>(s, none, p.rows, p.n, ws, d.C.data, d.alpha, beta_eff, splits, d);
        if (rc) return rc;
        EB_CUDA(cudaFreeAsync(ws, s));
    }
    return EB200_OK;
}

// ---------------------------------------------------------------- elementwise
template <typename T> int run_elementwise(cudaStream_t s, const Plan &p, eb200_contract_desc &d, const double beta_eff[2]) {
    ElemArgs a;
    a.t = p.t;
    a.A = d.A.data; a.B = d.B.data; a.C = d.C.data;
    a.alpha[0] = d.alpha[0]; a.alpha[1] = d.alpha[1];
    a.beta[0] = beta_eff[0]; a.beta[1] = beta_eff[1];
    int rc = k_launch_elem(Scalar<T>::dtype, s, a, d.conj_a, d.conj_b, grid_for(a.t.tot, 256, 16));
    if (rc) return rc;
    d.launches++;
    return EB200_OK;
}

// Whole-C group with C's own axes (every axis separately, repeated indices included): used when
// C = beta*C has to touch elements the contraction does not enumerate (C with a repeated index).
int full_c_group(const eb200_contract_desc &d, Grp &g) {
    memset(&g, 0, sizeof(g));
    g.tot = 1;
    // sort axes by stride, merge contiguous
    struct Ax { int64_t ext, st; };
    std::vector<Ax> ax;
    for (int a = 0; a < d.C.rank; a++) {
        int64_t e = d.extent[d.C.index[a]];
        if (e == 1) continue;
        ax.push_back({e, d.C.stride[a]});
    }
    std::stable_sort(ax.begin(), ax.end(), [](const Ax &x, const Ax &y) { return std::llabs(x.st) < std::llabs(y.st); });
    for (size_t i = 0; i + 1 < ax.size();) {
        if (ax[i + 1].st == ax[i].st * ax[i].ext) {
            ax[i].ext *= ax[i + 1].ext;
            ax.erase(ax.begin() + i + 1);
        } else i++;
    }
    if ((int)ax.size() > MAXG) return fail(EB200_ERR_UNSUPPORTED, "contract: C has more than %d non-mergeable axes", MAXG);
    for (auto &x : ax) {
        g.ext[g.n] = x.ext;
        g.s[ROLE_C][g.n] = x.st;
        g.n++;
        g.tot *= x.ext;
    }
    grp_finalize(g);
    return EB200_OK;
}

template <typename T> int contract_typed(cudaStream_t s, eb200_contract_desc &d) {
    Plan p;
    int  rc = make_plan(d, p, sm_count());
    if (rc) return rc;
    d.kernel_used = p.family;
    d.launches = 0;
    if (p.family == EB200_KF_NONE) return EB200_OK;
    if (!d.C.data) return fail(EB200_ERR_INVALID_ARGUMENT, "contract: C.data is NULL");

    double beta_eff[2] = {d.beta[0], Scalar<T>::is_complex ? d.beta[1] : 0.0};
    const bool beta_is_one = (beta_eff[0] == 1.0 && beta_eff[1] == 0.0);
    if (p.family == EB200_KF_SCALE_ONLY || p.c_has_repeat) {
        // reference semantics: C->zero() / *C *= beta on the WHOLE tensor (GPUTensorAlgebra.hpp:273-277)
        if (!beta_is_one) {
            Grp g;
            rc = full_c_group(d, g);
            if (rc) return rc;
            rc = launch_scale<T>(s, g, d.C.data, beta_eff, d);
            if (rc) return rc;
        }
        if (p.family == EB200_KF_SCALE_ONLY) return EB200_OK;
        beta_eff[0] = 1.0;
        beta_eff[1] = 0.0;
    }
    if (!d.A.data || !d.B.data) return fail(EB200_ERR_INVALID_ARGUMENT, "contract: A.data or B.data is NULL");
    switch (p.family) {
    case EB200_KF_TILE_MMA:
    case EB200_KF_TILE_SIMT: return run_tile<T>(s, p, d, beta_eff);
    case EB200_KF_ROWREDUCE:
    case EB200_KF_COLREDUCE:
    case EB200_KF_FULLREDUCE: return run_reduce<T>(s, p, d, beta_eff);
    case EB200_KF_ELEMENTWISE: return run_elementwise<T>(s, p, d, beta_eff);
    default: return fail(EB200_ERR_UNSUPPORTED, "contract: no kernel family for this descriptor");
    }
}

// ---------------------------------------------------------------- mixed element types
// The reference sends contractions whose tensors differ in value type (d<-f,d; f<-d,d; cd<-cd,d ...) to its generic kernel
// (Dispatch.hpp:896-899) and multiplies in the wider type. Here the narrower operands are widened into dense scratch
// copies (stream-ordered allocations) and the same-type engine runs in the promoted type; a narrower C is widened,
// updated and narrowed back. Rare path (tests and MixedPrecision-style calls), two small extra kernels.
size_t dtype_size(int dt) { return dt == EB200_F32 ? 4 : (dt == EB200_C128 ? 16 : 8); }

// dense row-major strides over the tensor's own axes; group for a strided<->dense copy
int axes_group(const eb200_contract_desc &d, const eb200_tensor_desc &t, int64_t *dense, Grp &g, bool to_dense) {
    memset(&g, 0, sizeof(g));
    g.tot = 1;
    if (t.rank > MAXG) return fail(EB200_ERR_UNSUPPORTED, "contract: mixed-type tensors of rank > %d are not supported", MAXG);
    int64_t acc = 1;
    for (int a = t.rank - 1; a >= 0; a--) {
        dense[a] = acc;
        acc *= d.extent[t.index[a]];
    }
    for (int a = t.rank - 1; a >= 0; a--) { // inner-first
        int i = g.n++;
        g.ext[i] = d.extent[t.index[a]];
        g.s[ROLE_A][i] = to_dense ? t.stride[a] : dense[a];
        g.s[ROLE_C][i] = to_dense ? dense[a] : t.stride[a];
        g.tot *= g.ext[i];
    }
    grp_finalize(g);
    return EB200_OK;
}

int contract_mixed(cudaStream_t s, eb200_contract_desc &d) {
    const int  dts[3] = {d.C.dtype, d.A.dtype, d.B.dtype};
    const bool any_c = dts[0] >= EB200_C64 || dts[1] >= EB200_C64 || dts[2] >= EB200_C64;
    const bool any_d = (dts[0] & 1) || (dts[1] & 1) || (dts[2] & 1); // F64 = 1, C128 = 3
    const int  W = any_c ? (any_d ? EB200_C128 : EB200_C64) : (any_d ? EB200_F64 : EB200_F32);
    eb200_contract_desc w = d;
    void               *tmp[3] = {nullptr, nullptr, nullptr};
    eb200_tensor_desc  *ts[3] = {&w.C, &w.A, &w.B};
    const eb200_tensor_desc *os[3] = {&d.C, &d.A, &d.B};
    Grp                 back; // dense -> strided for C
    int                 rc = EB200_OK;
    for (int i = 0; i < 3 && rc == EB200_OK; i++) {
        if (os[i]->dtype == W) continue;
        int64_t dense[EB200_MAX_RANK];
        Grp     g;
        if ((rc = axes_group(d, *os[i], dense, g, true))) break;
        if (g.tot == 0) {
            // an empty tensor has nothing to widen, but it must still take the promoted type: leaving its dtype alone sent the
            // descriptor straight back into contract_mixed (endless recursion, stack overflow)
            ts[i]->dtype = W;
            for (int a = 0; a < os[i]->rank; a++) ts[i]->stride[a] = dense[a];
            continue;
        }
        if ((rc = scratch_alloc(&tmp[i], (size_t)g.tot * dtype_size(W), s))) break;
        const bool need_values = (i != 0) || !(d.beta[0] == 0.0 && d.beta[1] == 0.0);
        if (need_values) {
            ConvArgs ca{g, os[i]->data, tmp[i]};
            rc = k_launch_convert(s, os[i]->dtype, W, ca, grid_for(ca.g.tot, 256));
            d.launches++;
        }
        ts[i]->data = tmp[i];
        ts[i]->dtype = W;
        for (int a = 0; a < os[i]->rank; a++) ts[i]->stride[a] = dense[a];
        if (i == 0) axes_group(d, *os[i], dense, back, false);
    }
    if (rc == EB200_OK) {
        int l0 = d.launches;
        if (w.A.dtype != w.C.dtype || w.B.dtype != w.C.dtype) rc = fail(EB200_ERR_UNSUPPORTED, "contract: mixed-type promotion did not converge");
        else rc = contract_dispatch(s, w);
        d.kernel_used = w.kernel_used;
        d.split_k = w.split_k;
        d.launches = l0 + w.launches;
    }
    if (rc == EB200_OK && tmp[0]) {
        ConvArgs ca{back, tmp[0], d.C.data};
        rc = k_launch_convert(s, W, d.C.dtype, ca, grid_for(ca.g.tot, 256));
        d.launches++;
    }
    for (int i = 0; i < 3; i++)
        if (tmp[i]) cudaFreeAsync(tmp[i], s);
    return rc;
}

} // namespace

int contract_dispatch(cudaStream_t s, eb200_contract_desc &d) {
    if (d.A.dtype != d.C.dtype || d.B.dtype != d.C.dtype) {
        for (int dt : {d.C.dtype, d.A.dtype, d.B.dtype})
            if (dt < EB200_F32 || dt > EB200_C128) return fail(EB200_ERR_INVALID_ARGUMENT, "contract: bad dtype %d", dt);
        if (d.n_unique < 0 || d.n_unique > EB200_MAX_UNIQUE) return fail(EB200_ERR_INVALID_ARGUMENT, "contract: n_unique out of range");
        d.launches = 0;
        return contract_mixed(s, d);
    }
    switch (d.C.dtype) {
    case EB200_F32: return contract_typed<float>(s, d);
    case EB200_F64: return contract_typed<double>(s, d);
    case EB200_C64: return contract_typed<float2>(s, d);
    case EB200_C128: return contract_typed<double2>(s, d);
    default: return fail(EB200_ERR_INVALID_ARGUMENT, "contract: bad dtype %d", d.C.dtype);
    }
}

} // namespace eb200

extern "C" {

int eb200_contract(void *stream, eb200_contract_desc *desc) {
    if (!desc) return eb200::fail(EB200_ERR_INVALID_ARGUMENT, "eb200_contract: desc is NULL");
    return eb200::contract_dispatch(eb200::as_stream(stream), *desc);
}

} // extern "C" (helpers below need C++ linkage)

namespace eb200 {
// Auxiliary streams of eb200_contract_grouped: per host thread, created on first use for the current device.
struct GroupStreams {
    static constexpr int N = 3;
    int          device = -1;
    cudaStream_t aux[N] = {};
    cudaEvent_t  fork = nullptr, join[N] = {};
    int ensure() {
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (dev == device) return N;
        release();
        for (int i = 0; i < N; i++)
            if (cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) != cudaSuccess) {
                release();
                return 0;
            }
        if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) {
            release();
            return 0;
        }
        device = dev;
        return N;
    }
    void release() {
        for (int i = 0; i < N; i++) {
            if (aux[i]) cudaStreamDestroy(aux[i]);
            if (join[i]) cudaEventDestroy(join[i]);
            aux[i] = nullptr; join[i] = nullptr;
        }
        if (fork) cudaEventDestroy(fork);
        fork = nullptr;
        device = -1;
    }
};
static GroupStreams &group_streams() {
    static thread_local GroupStreams gs; // not destroyed at thread exit on purpose: the context may already be gone
    return gs;
}
} // namespace eb200

// Grouped contraction (the per-block loops of TA/Backends/BlockAlgebra.hpp:36-47 and TileAlgebra.hpp:57-90).
// Neighbouring descriptors that differ only by base pointers advancing by a constant element step are folded into ONE
// contraction by giving them an extra unique index whose stride is that step:
//   C, A, B all step (or one operand is shared)  -> a batch / M / N index: equal-size blocks become one batched launch;
//   C is the same, A and B step, beta == 1       -> a contracted index: the accumulation over tile columns becomes K.
// Folding is repeated on the folded list, so a full grid of equal tiles collapses level by level (k, then j, then i) into
// a single launch. Whatever does not fold is spread over a few auxiliary streams forked from, and joined back into, the
// caller's stream; members that write the same C stay on one stream, in order. Members with different C must be
// independent of each other (distinct outputs, no output aliasing another member's input), as blocks and tiles are.
namespace eb200 {
struct GroupItem {
    eb200_contract_desc d;
    int64_t             first, count; // range of the caller's descriptors this item covers
};
static bool group_same_shape(const eb200_contract_desc &x, const eb200_contract_desc &y) {
    if (x.n_unique != y.n_unique || x.conj_a != y.conj_a || x.conj_b != y.conj_b || x.force_kernel != y.force_kernel || x.split_k != y.split_k) return false;
    for (int k = 0; k < 2; k++)
        if (x.alpha[k] != y.alpha[k] || x.beta[k] != y.beta[k]) return false;
    for (int u = 0; u < x.n_unique; u++)
        if (x.extent[u] != y.extent[u]) return false;
    const eb200_tensor_desc *tx[3] = {&x.C, &x.A, &x.B}, *ty[3] = {&y.C, &y.A, &y.B};
    for (int t = 0; t < 3; t++) {
        if (tx[t]->dtype != ty[t]->dtype || tx[t]->rank != ty[t]->rank) return false;
        for (int i = 0; i < tx[t]->rank; i++)
            if (tx[t]->index[i] != ty[t]->index[i] || tx[t]->stride[i] != ty[t]->stride[i]) return false;
    }
    return true;
}
static bool group_steps(const eb200_contract_desc &x, const eb200_contract_desc &y, int64_t st[3]) { // element steps y - x of C, A, B
    const eb200_tensor_desc *tx[3] = {&x.C, &x.A, &x.B}, *ty[3] = {&y.C, &y.A, &y.B};
    for (int t = 0; t < 3; t++) {
        ptrdiff_t d = (const char *)ty[t]->data - (const char *)tx[t]->data, es = (ptrdiff_t)dtype_size(tx[t]->dtype);
        if (d % es) return false;
        st[t] = d / es;
    }
    return true;
}
static bool group_foldable(const eb200_contract_desc &d, const int64_t st[3]) {
    if (st[0] > 0) return st[1] >= 0 && st[2] >= 0 && (st[1] > 0 || st[2] > 0);
    if (st[0] == 0) return st[1] > 0 && st[2] > 0 && d.beta[0] == 1.0 && d.beta[1] == 0.0; // accumulation into one C
    return false;
}
/// one folding pass over `n` descriptors reached through `at(i)`; emits one item per run
template <typename At, typename Range> static void group_fold_pass(size_t n, At &&at, Range &&range_of, std::vector<GroupItem> &out) {
    out.clear();
    for (size_t i = 0; i < n;) {
        const eb200_contract_desc &first = at(i);
        size_t                     cnt = 1;
        int64_t                    st[3] = {0, 0, 0};
        const bool room = first.n_unique < EB200_MAX_UNIQUE && first.C.rank < EB200_MAX_RANK && first.A.rank < EB200_MAX_RANK && first.B.rank < EB200_MAX_RANK;
        if (room && i + 1 < n && group_same_shape(first, at(i + 1)) && group_steps(first, at(i + 1), st) && group_foldable(first, st)) {
            cnt = 2;
            int64_t nx[3];
            while (i + cnt < n && group_same_shape(first, at(i + cnt)) && group_steps(at(i + cnt - 1), at(i + cnt), nx) && nx[0] == st[0] &&
                   nx[1] == st[1] && nx[2] == st[2])
                cnt++;
        }
        out.emplace_back();
        GroupItem &m = out.back();
        m.d = first;
        m.first = range_of(i).first;
        m.count = 0;
        for (size_t k = 0; k < cnt; k++) m.count += range_of(i + k).second;
        if (cnt > 1) {
            const int g = m.d.n_unique++;
            m.d.extent[g] = (int64_t)cnt;
            eb200_tensor_desc *t3[3] = {&m.d.C, &m.d.A, &m.d.B};
            for (int t = 0; t < 3; t++) {
                if (st[t] == 0) continue; // shared by every member: the new index does not address it
                eb200_tensor_desc &td = *t3[t];
                for (int k = td.rank; k > 0; k--) { td.index[k] = td.index[k - 1]; td.stride[k] = td.stride[k - 1]; }
                td.index[0] = g;
                td.stride[0] = st[t];
                td.rank++;
            }
        }
        i += cnt;
    }
}
} // namespace eb200

extern "C" {

int eb200_contract_grouped(void *stream, int64_t n, eb200_contract_desc *descs, int32_t *n_launched_groups) {
    using namespace eb200;
    if (n < 0 || (n > 0 && !descs)) return fail(EB200_ERR_INVALID_ARGUMENT, "eb200_contract_grouped: bad arguments");
    std::vector<GroupItem> items, next;
    // level 1 reads the caller's array in place; further levels fold the folded list until nothing changes
    group_fold_pass((size_t)n, [&](size_t i) -> const eb200_contract_desc & { return descs[i]; },
                    [&](size_t i) { return std::pair<int64_t, int64_t>((int64_t)i, 1); }, items);
    for (int pass = 0; pass < 6 && items.size() > 1; pass++) {
        group_fold_pass(items.size(), [&](size_t i) -> const eb200_contract_desc & { return items[i].d; },
                        [&](size_t i) { return std::pair<int64_t, int64_t>(items[i].first, items[i].count); }, next);
        const bool shrunk = next.size() < items.size();
        items.swap(next);
        if (!shrunk) break;
    }

    cudaStream_t  user = as_stream(stream);
    GroupStreams &gs = group_streams();
    const int     naux = items.size() > 1 ? gs.ensure() : 0;
    if (naux > 0) {
        EB_CUDA(cudaEventRecord(gs.fork, user));
        for (int a = 0; a < naux; a++) EB_CUDA(cudaStreamWaitEvent(gs.aux[a], gs.fork, 0));
    }
    // Lane assignment by the byte range of C each (possibly folded) item touches: items whose ranges overlap share a lane
    // and therefore stay ordered, also when a folded item covers C0..Cn and a later one accumulates into Ck (k > 0).
    // Ranges are kept as disjoint intervals (lo -> hi, lane); an item overlapping intervals of different lanes forces a
    // join + re-fork of all lanes first (rare).
    struct Span { uintptr_t hi; int lane; };
    std::map<uintptr_t, Span> spans;
    auto c_range = [](const eb200_contract_desc &d, uintptr_t &lo, uintptr_t &hi) {
        const int64_t es = (int64_t)dtype_size(d.C.dtype);
        int64_t       mn = 0, mx = 0;
        for (int a = 0; a < d.C.rank; a++) {
            const int64_t e = d.extent[d.C.index[a]];
            if (e <= 0) { lo = hi = (uintptr_t)d.C.data; return; }
            const int64_t reach = (e - 1) * d.C.stride[a];
            if (reach < 0) mn += reach; else mx += reach;
        }
        lo = (uintptr_t)((const char *)d.C.data + mn * es);
        hi = (uintptr_t)((const char *)d.C.data + (mx + 1) * es);
    };
    int rc = EB200_OK;
    for (size_t r = 0; r < items.size() && rc == EB200_OK; r++) {
        GroupItem &it = items[r];
        int        lane = 0;
        if (naux > 0) {
            uintptr_t lo, hi;
            c_range(it.d, lo, hi);
            lane = int(((lo / 256u) * 2654435761u >> 16) % (unsigned)(naux + 1));
            if (hi > lo) {
                auto first = spans.upper_bound(lo);
                if (first != spans.begin() && std::prev(first)->second.hi > lo) --first;
                auto last = first;
                int  seen = -1;
                bool conflict = false;
                for (; last != spans.end() && last->first < hi; ++last) {
                    if (seen >= 0 && last->second.lane != seen) conflict = true;
                    seen = last->second.lane;
                    lo = std::min(lo, last->first);
                    hi = std::max(hi, last->second.hi);
                }
                if (conflict) { // order this item after everything enqueued so far, then carry on in parallel
                    for (int a = 0; a < naux; a++) {
                        cudaEventRecord(gs.join[a], gs.aux[a]);
                        cudaStreamWaitEvent(user, gs.join[a], 0);
                    }
                    cudaEventRecord(gs.fork, user);
                    for (int a = 0; a < naux; a++) cudaStreamWaitEvent(gs.aux[a], gs.fork, 0);
                    spans.clear();
                    first = last = spans.end();
                    lane = 0;
                } else if (seen >= 0) lane = seen;
                spans.erase(first, last);
                spans[lo] = Span{hi, lane};
            }
        }
        cudaStream_t s = lane == 0 ? user : gs.aux[lane - 1];
        rc = contract_dispatch(s, it.d);
        for (int64_t k = 0; k < it.count; k++) {
            descs[it.first + k].kernel_used = it.d.kernel_used;
            descs[it.first + k].launches = k == 0 ? it.d.launches : 0;
        }
    }
    if (naux > 0) { // always join, also after an error, so the caller's stream never runs ahead of forked work
        for (int a = 0; a < naux; a++) {
            cudaEventRecord(gs.join[a], gs.aux[a]);
            cudaStreamWaitEvent(user, gs.join[a], 0);
        }
    }
    if (rc) return rc;
    if (n_launched_groups) *n_launched_groups = (int32_t)items.size();
    return EB200_OK;
}

int eb200_contract_plan(eb200_contract_desc *desc) {
    if (!desc) return eb200::fail(EB200_ERR_INVALID_ARGUMENT, "eb200_contract_plan: desc is NULL");
    eb200::Plan p;
    int         rc = eb200::make_plan(*desc, p, 148);
    if (rc) return rc;
    desc->kernel_used = p.family;
    return EB200_OK;
}

} // extern "C"
