// plan.cu — host planner for eb200_contract (see plan.cuh). Pure host code.
#include "plan.cuh"

#include <algorithm>
#include <climits>
#include <vector>

namespace eb200 {
namespace {

struct Idx {
    int64_t ext;
    int64_t s[3];   // summed stride per tensor (Hadamard: an index repeated inside one tensor adds its strides,
                    // exactly what the reference's get_stride_for does, TensorAlgebra/Detail/Utilities.hpp:560-592)
    bool    in[3];
};

inline void grp_clear(Grp &g) {
    memset(&g, 0, sizeof(g));
    g.tot = 1;
    grp_finalize(g);
}

inline bool grp_push(Grp &g, const Idx &x) {
    if (g.n >= MAXG) return false;
    g.ext[g.n] = x.ext;
    for (int r = 0; r < 3; r++) g.s[r][g.n] = x.s[r];
    g.n++;
    g.tot *= x.ext;
    return true;
}

inline int64_t key_of(int64_t stride) { return stride == 0 ? INT64_MAX : (stride < 0 ? -stride : stride); }

// order inner-first by |stride| in tensor `r0`, ties by tensor `r1`; zero strides (broadcast) go outermost
void sort_items(std::vector<Idx> &v, int r0, int r1) {
    std::stable_sort(v.begin(), v.end(), [&](const Idx &a, const Idx &b) {
        int64_t ka = key_of(a.s[r0]), kb = key_of(b.s[r0]);
        if (ka != kb) return ka < kb;
        return key_of(a.s[r1]) < key_of(b.s[r1]);
    });
}

// fuse neighbours (inner i, outer i+1) that are contiguous in every tensor
void merge_items(std::vector<Idx> &v) {
    size_t i = 0;
    while (i + 1 < v.size()) {
        bool ok = true;
        for (int r = 0; r < 3; r++)
            if (v[i + 1].s[r] != v[i].s[r] * v[i].ext) ok = false;
        if (ok) {
            v[i].ext *= v[i + 1].ext;
            v.erase(v.begin() + i + 1);
        } else {
            i++;
        }
    }
}

bool build(Grp &g, std::vector<Idx> v, int r0, int r1) {
    sort_items(v, r0, r1);
    merge_items(v);
    grp_clear(g);
    for (auto &x : v)
        if (!grp_push(g, x)) return false;
    grp_finalize(g);
    return true;
}

int64_t min_stride(const std::vector<Idx> &v, int role) {
    int64_t m = INT64_MAX;
    for (auto &x : v) m = std::min(m, key_of(x.s[role]));
    return m;
}

} // namespace

int make_plan(const eb200_contract_desc &d, Plan &p, int /*sm_count_hint*/) {
    memset(&p, 0, sizeof(p));
    if (d.n_unique < 0 || d.n_unique > EB200_MAX_UNIQUE)
        return fail(EB200_ERR_INVALID_ARGUMENT, "contract: n_unique=%d outside [0,%d]", d.n_unique, EB200_MAX_UNIQUE);
    const eb200_tensor_desc *ts[3] = {&d.A, &d.B, &d.C};
    const char              *nm[3] = {"A", "B", "C"};
    for (int r = 0; r < 3; r++) {
        if (ts[r]->rank < 0 || ts[r]->rank > EB200_MAX_RANK)
            return fail(EB200_ERR_RANK, "contract: rank of %s is %d, outside [0,%d]", nm[r], ts[r]->rank, EB200_MAX_RANK);
        for (int a = 0; a < ts[r]->rank; a++)
            if (ts[r]->index[a] < 0 || ts[r]->index[a] >= d.n_unique)
                return fail(EB200_ERR_INVALID_ARGUMENT, "contract: %s.index[%d]=%d is not a valid unique-index position (n_unique=%d)", nm[r], a,
                            ts[r]->index[a], d.n_unique);
        if (ts[r]->dtype < EB200_F32 || ts[r]->dtype > EB200_C128)
            return fail(EB200_ERR_INVALID_ARGUMENT, "contract: %s.dtype=%d is not an eb200_dtype", nm[r], ts[r]->dtype);
    }
    for (int u = 0; u < d.n_unique; u++)
        if (d.extent[u] < 0) return fail(EB200_ERR_DOMAIN, "contract: extent[%d]=%lld is negative", u, (long long)d.extent[u]);

    std::vector<Idx> idx(d.n_unique);
    int              count_in_c[EB200_MAX_UNIQUE] = {0};
    for (int u = 0; u < d.n_unique; u++) {
        idx[u].ext = d.extent[u];
        for (int r = 0; r < 3; r++) {
            idx[u].s[r]  = 0;
            idx[u].in[r] = false;
        }
    }
    for (int r = 0; r < 3; r++)
        for (int a = 0; a < ts[r]->rank; a++) {
            int u = ts[r]->index[a];
            idx[u].s[r] += ts[r]->stride[a];
            idx[u].in[r] = true;
            if (r == ROLE_C) count_in_c[u]++;
        }
    for (int u = 0; u < d.n_unique; u++)
        if (count_in_c[u] > 1) p.c_has_repeat = 1;

    // empty spaces. An index nobody carries is ignored (extent irrelevant).
    for (int u = 0; u < d.n_unique; u++) {
        if (!(idx[u].in[0] || idx[u].in[1] || idx[u].in[2])) continue;
        if (idx[u].ext == 0) {
            if (idx[u].in[ROLE_C]) p.empty_c = 1;
            else p.empty_k = 1;
        }
    }
    grp_clear(p.b); grp_clear(p.m); grp_clear(p.n); grp_clear(p.k); grp_clear(p.t); grp_clear(p.rows);
    if (p.empty_c) {
        p.family = EB200_KF_NONE;
        return EB200_OK;
    }

    std::vector<Idx> vb, vm, vn, vk, vt;
    for (int u = 0; u < d.n_unique; u++) {
        const Idx &x = idx[u];
        if (!(x.in[0] || x.in[1] || x.in[2])) continue;
        if (x.in[ROLE_C] && x.ext >= 1) {
            // every C index (even extent 1) is kept in the scaling/elementwise group only if extent > 1
        }
        if (x.ext == 1) continue; // extent-1 indices contribute offset 0 everywhere
        if (x.ext == 0) continue; // handled through empty_k
        if (x.in[ROLE_C]) {
            vt.push_back(x);
            if (x.in[ROLE_A] && x.in[ROLE_B]) vb.push_back(x);
            else if (x.in[ROLE_B]) vn.push_back(x);
            else vm.push_back(x); // (C,A) targets and C-only broadcast indices (A stride 0)
        } else {
            vk.push_back(x); // links, plus indices summed in A only / B only (other stride 0)
        }
    }

    if (!build(p.t, vt, ROLE_C, ROLE_A)) return fail(EB200_ERR_UNSUPPORTED, "contract: more than %d non-mergeable target indices", MAXG);

    if (p.empty_k) {
        p.family = EB200_KF_SCALE_ONLY;
        return EB200_OK;
    }

    const int64_t Mtot = [&] { int64_t t = 1; for (auto &x : vm) t *= x.ext; return t; }();
    const int64_t Ntot = [&] { int64_t t = 1; for (auto &x : vn) t *= x.ext; return t; }();

    int family;
    if (vk.empty()) family = EB200_KF_ELEMENTWISE;
    else if (std::min(Mtot, Ntot) <= 8) family = EB200_KF_ROWREDUCE; // refined below
    else family = (d.C.dtype == EB200_F64) ? EB200_KF_TILE_MMA : EB200_KF_TILE_SIMT;

    if (d.force_kernel == EB200_KF_TILE_MMA || d.force_kernel == EB200_KF_TILE_SIMT) {
        if (!vk.empty()) family = (d.force_kernel == EB200_KF_TILE_MMA && d.C.dtype == EB200_F64) ? EB200_KF_TILE_MMA : EB200_KF_TILE_SIMT;
    } else if (d.force_kernel == EB200_KF_ROWREDUCE || d.force_kernel == EB200_KF_COLREDUCE) {
        if (!vk.empty() && std::min(Mtot, Ntot) <= 8) family = d.force_kernel;
    }

    if (family == EB200_KF_ROWREDUCE || family == EB200_KF_COLREDUCE) {
        // make B the operand with the few columns: exchange A<->B and M<->N when M is the small side
        if (Mtot < Ntot || (Mtot == Ntot && false)) {
            p.swapped = 1;
            std::swap(vm, vn);
            auto swap_roles = [](std::vector<Idx> &v) {
                for (auto &x : v) {
                    std::swap(x.s[ROLE_A], x.s[ROLE_B]);
                    std::swap(x.in[ROLE_A], x.in[ROLE_B]);
                }
            };
            swap_roles(vm); swap_roles(vn); swap_roles(vk); swap_roles(vb);
        }
        std::vector<Idx> vrows = vb;
        vrows.insert(vrows.end(), vm.begin(), vm.end());
        if (d.force_kernel != EB200_KF_ROWREDUCE && d.force_kernel != EB200_KF_COLREDUCE) {
            int64_t mk = min_stride(vk, ROLE_A), mr = min_stride(vrows, ROLE_A);
            family = (mk <= mr) ? EB200_KF_ROWREDUCE : EB200_KF_COLREDUCE;
        }
        if (!build(p.rows, vrows, ROLE_A, ROLE_C) || !build(p.n, vn, ROLE_C, ROLE_B) || !build(p.k, vk, ROLE_A, ROLE_B))
            return fail(EB200_ERR_UNSUPPORTED, "contract: more than %d non-mergeable indices in one group", MAXG);
        // dot-like with everything contracted: rows.tot == 1
        if (p.rows.tot == 1 && p.n.tot == 1 && d.force_kernel == 0) family = EB200_KF_FULLREDUCE;
    } else if (family == EB200_KF_TILE_MMA || family == EB200_KF_TILE_SIMT) {
        p.a_kmajor = min_stride(vk, ROLE_A) <= min_stride(vm, ROLE_A);
        p.b_kmajor = min_stride(vk, ROLE_B) <= min_stride(vn, ROLE_B);
        if (!build(p.b, vb, ROLE_C, ROLE_A) || !build(p.m, vm, ROLE_A, ROLE_C) || !build(p.n, vn, ROLE_B, ROLE_C) ||
            !build(p.k, vk, p.a_kmajor ? ROLE_A : ROLE_B, p.a_kmajor ? ROLE_B : ROLE_A))
            return fail(EB200_ERR_UNSUPPORTED, "contract: more than %d non-mergeable indices in one group", MAXG);
    }
    p.family = family;
    return EB200_OK;
}

} // namespace eb200
