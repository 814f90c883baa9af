// gs.cuh -- gather-scatter (direct-stiffness summation).
//
// Replaces gslib v1.0.9 gs_setup / gs_op as driven by core/dssum.f:1-31 (setupds), :33-98 (dssum),
// :100-161 (dsop), :163-208 (vec_dssum -> gs_op_many), :260-287 (nvec_dssum -> gs_op_fields).
// Semantics: all entries (local and remote) carrying one non-zero id are replaced by their
// sum / product / min / max; id 0 entries are untouched.
//
// Setup (device, CUB): compact non-zero ids -> stable radix sort by id -> run-length encode ->
// keep runs of length >= 2 -> order groups by their first member -> CSR (goff, gidx).
// Members are ascending inside a group and combined in that order (deterministic, no atomics).
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "ctx.cuh"

namespace nekb {

struct NonZeroId {
    const int64_t *id;
    __host__ __device__ bool operator()(const int32_t &i) const { return id[i] != 0; }
};
struct CountGe2 {
    const int32_t *cnt;
    __host__ __device__ bool operator()(const int32_t &r) const { return cnt[r] >= 2; }
};

__global__ void gs_gather_keys(int64_t *keys, const int64_t *id, const int32_t *idx, int64_t m)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < m; t += (int64_t)gridDim.x * blockDim.x)
        keys[t] = id[idx[t]];
}
__global__ void gs_group_first(int32_t *first, int32_t *gcnt, const int32_t *sel_run, const int32_t *run_off,
                               const int32_t *run_cnt, const int32_t *sorted_idx, int64_t ng)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ng; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = sel_run[t];
        first[t] = sorted_idx[run_off[r]];
        gcnt[t] = run_cnt[r];
    }
}
__global__ void gs_permute_counts(int32_t *cnt_out, const int32_t *gcnt, const int32_t *order, int64_t ng)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ng; t += (int64_t)gridDim.x * blockDim.x)
        cnt_out[t] = gcnt[order[t]];
}
__global__ void gs_fill_members(int32_t *gidx, const int32_t *goff, const int32_t *order, const int32_t *sel_run,
                                const int32_t *run_off, const int32_t *sorted_idx, int64_t ng)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ng; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = sel_run[order[t]];
        const int b = goff[t], cnt = goff[t + 1] - b, src = run_off[r];
        for (int q = 0; q < cnt; q++) gidx[b + q] = sorted_idx[src + q];
    }
}

template <int OP>
__device__ __forceinline__ double gs_combine(double a, double b)
{
    if (OP == 1) return a + b;
    if (OP == 2) return a * b;
    if (OP == 3) return fmin(a, b);
    return fmax(a, b);
}

template <int OP>
__global__ void __launch_bounds__(256)
    gs_local_kernel(double *__restrict__ u, const int32_t *__restrict__ goff, const int32_t *__restrict__ gidx,
                    int ngroups)
{
    for (int gI = blockIdx.x * blockDim.x + threadIdx.x; gI < ngroups; gI += gridDim.x * blockDim.x) {
        const int b = goff[gI], e = goff[gI + 1];
        double v = u[gidx[b]];
        for (int q = b + 1; q < e; q++) v = gs_combine<OP>(v, u[gidx[q]]);
        for (int q = b; q < e; q++) u[gidx[q]] = v;
    }
}

__global__ void __launch_bounds__(256) col2_kernel(double *__restrict__ a, const double *__restrict__ b, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        a[t] *= b[t];
}

inline int blocks_for(int64_t n, int threads = 256)
{
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)ctx().num_sms * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// Builds the local map of `h` from device-resident ids.
inline void gs_build_local(GsMap &h, const int64_t *id_dev, int64_t n)
{
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    NEKB_REQUIRE(n < (int64_t)2147483647, "gs_setup: local vector too long for int32 indexing");
    h.n = n;
    h.ngroups = h.nmembers = 0;
    h.link_mode = 0;
    if (n == 0) return;

    DevBuf<int32_t> idx_nz, idx_sorted, num;
    DevBuf<int64_t> keys, keys_sorted;
    DevBuf<char> tmp;
    idx_nz.alloc(n);
    num.alloc(4);
    size_t tb = 0;
    thrust::counting_iterator<int32_t> iota(0);
    NonZeroId pred{id_dev};
    NEKB_CUDA(cub::DeviceSelect::If(nullptr, tb, iota, idx_nz.p, num.p, (int)n, pred, s));
    tmp.alloc(tb);
    NEKB_CUDA(cub::DeviceSelect::If(tmp.p, tb, iota, idx_nz.p, num.p, (int)n, pred, s));
    launch_counter() += 1;
    int m = 0;
    NEKB_CUDA(cudaMemcpyAsync(&m, num.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (m == 0) return;

    keys.alloc(m);
    keys_sorted.alloc(m);
    idx_sorted.alloc(m);
    gs_gather_keys<<<blocks_for(m), 256, 0, s>>>(keys.p, id_dev, idx_nz.p, m);
    NEKB_LAUNCHED();
    NEKB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys_sorted.p, idx_nz.p, idx_sorted.p, m, 0, 64, s));
    tmp.ensure(tb);
    NEKB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys_sorted.p, idx_nz.p, idx_sorted.p, m, 0, 64, s));
    launch_counter() += 1;
    keys.release();
    idx_nz.release();

    // run-length encode the sorted ids
    DevBuf<int64_t> uniq;
    DevBuf<int32_t> run_cnt, run_off;
    uniq.alloc(m);
    run_cnt.alloc(m);
    NEKB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb, keys_sorted.p, uniq.p, run_cnt.p, num.p, m, s));
    tmp.ensure(tb);
    NEKB_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tb, keys_sorted.p, uniq.p, run_cnt.p, num.p, m, s));
    int nruns = 0;
    NEKB_CUDA(cudaMemcpyAsync(&nruns, num.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    uniq.release();
    keys_sorted.release();
    run_off.alloc(nruns + 1);
    NEKB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, run_cnt.p, run_off.p, nruns, s));
    tmp.ensure(tb);
    NEKB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, run_cnt.p, run_off.p, nruns, s));

    // keep runs with >= 2 members
    DevBuf<int32_t> sel_run;
    sel_run.alloc(nruns);
    CountGe2 p2{run_cnt.p};
    NEKB_CUDA(cub::DeviceSelect::If(nullptr, tb, iota, sel_run.p, num.p, nruns, p2, s));
    tmp.ensure(tb);
    NEKB_CUDA(cub::DeviceSelect::If(tmp.p, tb, iota, sel_run.p, num.p, nruns, p2, s));
    int ng = 0;
    NEKB_CUDA(cudaMemcpyAsync(&ng, num.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (ng == 0) return;

    // order the groups by their first (smallest) member so neighbouring threads touch neighbouring memory
    DevBuf<int32_t> first, first_sorted, gcnt, order, order_in, cnt_perm;
    first.alloc(ng);
    first_sorted.alloc(ng);
    gcnt.alloc(ng);
    order.alloc(ng);
    order_in.alloc(ng);
    cnt_perm.alloc(ng);
    gs_group_first<<<blocks_for(ng), 256, 0, s>>>(first.p, gcnt.p, sel_run.p, run_off.p, run_cnt.p, idx_sorted.p, ng);
    NEKB_LAUNCHED();
    {
        std::vector<int32_t> io(ng);
        for (int t = 0; t < ng; t++) io[t] = t;
        NEKB_CUDA(cudaMemcpyAsync(order_in.p, io.data(), sizeof(int32_t) * ng, cudaMemcpyHostToDevice, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
    }
    NEKB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, first.p, first_sorted.p, order_in.p, order.p, ng, 0, 32, s));
    tmp.ensure(tb);
    NEKB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, first.p, first_sorted.p, order_in.p, order.p, ng, 0, 32, s));
    gs_permute_counts<<<blocks_for(ng), 256, 0, s>>>(cnt_perm.p, gcnt.p, order.p, ng);
    NEKB_LAUNCHED();
    h.goff.alloc((size_t)ng + 1);
    NEKB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt_perm.p, h.goff.p, ng, s));
    tmp.ensure(tb);
    NEKB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt_perm.p, h.goff.p, ng, s));
    // total members = last offset + last count
    int last_off = 0, last_cnt = 0;
    NEKB_CUDA(cudaMemcpyAsync(&last_off, h.goff.p + (ng - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaMemcpyAsync(&last_cnt, cnt_perm.p + (ng - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    const int nm = last_off + last_cnt;
    NEKB_CUDA(cudaMemcpyAsync(h.goff.p + ng, &nm, sizeof(int), cudaMemcpyHostToDevice, s));
    h.gidx.alloc(nm);
    gs_fill_members<<<blocks_for(ng), 256, 0, s>>>(h.gidx.p, h.goff.p, order.p, sel_run.p, run_off.p, idx_sorted.p, ng);
    NEKB_LAUNCHED();
    NEKB_CUDA(cudaStreamSynchronize(s));
    h.ngroups = ng;
    h.nmembers = nm;
}

inline int gs_new_handle()
{
    Ctx &c = ctx();
    for (size_t i = 0; i < c.gs.size(); i++)
        if (!c.gs[i].used) {
            c.gs[i] = GsMap();
            c.gs[i].used = true;
            return (int)i;
        }
    c.gs.emplace_back();
    c.gs.back().used = true;
    return (int)c.gs.size() - 1;
}

inline GsMap &gs_get(int handle)
{
    Ctx &c = ctx();
    NEKB_REQUIRE(handle >= 0 && handle < (int)c.gs.size() && c.gs[handle].used, "invalid gs handle");
    return c.gs[handle];
}

inline void gs_remote_exchange(GsMap &h, double *u, int op);  // comm.cuh

inline void gs_local(GsMap &h, double *u, int op)
{
    Ctx &c = ctx();
    if (h.ngroups == 0) return;
    const int grid = blocks_for(h.ngroups);
    switch (op) {
        case 1: gs_local_kernel<1><<<grid, 256, 0, c.stream>>>(u, h.goff.p, h.gidx.p, (int)h.ngroups); break;
        case 2: gs_local_kernel<2><<<grid, 256, 0, c.stream>>>(u, h.goff.p, h.gidx.p, (int)h.ngroups); break;
        case 3: gs_local_kernel<3><<<grid, 256, 0, c.stream>>>(u, h.goff.p, h.gidx.p, (int)h.ngroups); break;
        case 4: gs_local_kernel<4><<<grid, 256, 0, c.stream>>>(u, h.goff.p, h.gidx.p, (int)h.ngroups); break;
        default: NEKB_REQUIRE(false, "gs_op: unsupported op (1 +, 2 *, 3 min, 4 max)");
    }
    NEKB_LAUNCHED();
}

// Node -> group view of the same map, for kernels that GATHER the assembled value of a node while streaming over the
// vector (cg.cuh cggos_update2_gs_kernel) instead of scattering sums back: face nodes (pairs) hold the partner's index,
// edge / corner nodes the group number.  Built on first use, 4 bytes per node.
__global__ void __launch_bounds__(256)
    gs_build_link_kernel(int32_t *__restrict__ link, const int32_t *__restrict__ goff, const int32_t *__restrict__ gidx,
                         int ngroups, int pairs_only)
{
    for (int gI = blockIdx.x * blockDim.x + threadIdx.x; gI < ngroups; gI += gridDim.x * blockDim.x) {
        const int b = goff[gI], e = goff[gI + 1];
        if (e - b == 2) {
            link[gidx[b]] = gidx[b + 1];
            link[gidx[b + 1]] = gidx[b];
        } else if (!pairs_only) {
            for (int q = b; q < e; q++) link[gidx[q]] = -2 - gI;
        }
    }
}

// mode 1: every group is reachable from `link`; mode 2: only pairs are, and the groups with three or more members (edge and
// corner nodes: 19 of an interior element's 127 groups) get their own CSR so that gs_local_kernel can assemble them in place
// before the gathering kernel runs (their members then read their own, already assembled, value).
inline void gs_ensure_link(GsMap &h, int mode)
{
    if (h.link_mode == mode) return;
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    h.link.alloc((size_t)(h.n > 0 ? h.n : 1));
    NEKB_CUDA(cudaMemsetAsync(h.link.p, 0xFF, sizeof(int32_t) * (size_t)(h.n > 0 ? h.n : 1), s));
    if (h.ngroups > 0) {
        gs_build_link_kernel<<<blocks_for(h.ngroups), 256, 0, s>>>(h.link.p, h.goff.p, h.gidx.p, (int)h.ngroups, mode == 2);
        NEKB_LAUNCHED();
    }
    h.ngroups3 = 0;
    if (mode == 2 && h.ngroups > 0) {   // set-up only: filtered on the host
        std::vector<int32_t> off((size_t)h.ngroups + 1), idx((size_t)h.nmembers), off3(1, 0), idx3;
        h.goff.download(off.data(), off.size(), s);
        h.gidx.download(idx.data(), idx.size(), s);
        for (int64_t g = 0; g < h.ngroups; g++)
            if (off[g + 1] - off[g] >= 3) {
                idx3.insert(idx3.end(), idx.begin() + off[g], idx.begin() + off[g + 1]);
                off3.push_back((int32_t)idx3.size());
            }
        h.ngroups3 = (int64_t)off3.size() - 1;
        if (idx3.empty()) idx3.push_back(0);
        h.goff3.upload(off3.data(), off3.size(), s);
        h.gidx3.upload(idx3.data(), idx3.size(), s);
        NEKB_CUDA(cudaStreamSynchronize(s));
    }
    h.link_mode = mode;
}

// The value gs_op(+) would leave at a node, read from the un-assembled vector: same members, same (ascending) order as
// gs_local_kernel<1>, hence the same bits.
__device__ __forceinline__ double gs_gathered(const double *__restrict__ u, double own, int l,
                                              const int32_t *__restrict__ goff, const int32_t *__restrict__ gidx)
{
    if (l == -1) return own;
    if (l >= 0) return own + u[l];          // a pair: IEEE addition commutes, so member order does not matter
    const int g = -2 - l;
    const int b = goff[g], e = goff[g + 1];
    double v = u[gidx[b]];
    for (int q = b + 1; q < e; q++) v += u[gidx[q]];
    return v;
}

// ---------------------------------------------------------------------------------------------- structured gather view
// (node classes / slot numbering of an 8^3 tile: gs_st_nb, gs_st_slot, gs_st_face in ctx.cuh)
// Builds the structured view (host, set-up only).  Classification of a group of the local map:
//   S  (assembled in place)   : a member is shared with another rank, or the group mixes node classes / is not a clean pair
//   P  (gathered through ftab): exactly two face-interior members, and all 36 pairs of both faces follow one affine map
//   E  (assembled into gval)  : every member is an edge / corner node
inline bool gs_ensure_struct(GsMap &h)
{
    if (h.struct_state != 0) return h.struct_state > 0;
    h.struct_state = -1;
    if (h.n <= 0 || h.n % GS_ST_N3 != 0) return false;
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    const int64_t nel = h.n / GS_ST_N3, ng = h.ngroups;
    std::vector<int32_t> off((size_t)ng + 1, 0), idx((size_t)(h.nmembers > 0 ? h.nmembers : 1));
    if (ng > 0) {
        h.goff.download(off.data(), off.size(), s);
        h.gidx.download(idx.data(), (size_t)h.nmembers, s);
    }
    std::vector<unsigned char> shared((size_t)h.n, 0);
    if (h.nshared > 0 && h.nx_members > 0) {
        std::vector<int32_t> xg((size_t)h.nx_members);
        h.x_gidx.download(xg.data(), xg.size(), s);
        for (int32_t q : xg) shared[q] = 1;
    }
    auto cls = [](int32_t node) {
        const int q = node & (GS_ST_N3 - 1);
        return gs_st_nb(q & 7, (q >> 3) & 7, q >> 6);
    };
    // pass 1: classify; partner[] of the candidate pairs
    enum : unsigned char { GS_S = 0, GS_P = 1, GS_E = 2 };
    std::vector<unsigned char> gclass((size_t)(ng > 0 ? ng : 1), GS_S);
    std::vector<int32_t> partner((size_t)h.n, -1), gof((size_t)h.n, -1);
    for (int64_t g = 0; g < ng; g++) {
        const int b = off[g], e = off[g + 1];
        bool sh = false, all_face = true, all_edge = true;
        for (int q = b; q < e; q++) {
            const int nb = cls(idx[q]);
            sh = sh || shared[idx[q]];
            all_face = all_face && nb == 1;
            all_edge = all_edge && nb >= 2;
            gof[idx[q]] = (int32_t)g;
        }
        if (sh) continue;
        if (e - b == 2 && all_face) {
            gclass[g] = GS_P;
            partner[idx[b]] = idx[b + 1], partner[idx[b + 1]] = idx[b];
        } else if (all_edge)
            gclass[g] = GS_E;
    }
    // pass 2: affine fit per face; a face that is only partly paired or not affine sends its groups to S
    std::vector<FaceLink> ftab((size_t)nel * 6);
    auto node_of = [](int f, int a, int b) {   // local index of in-face (a, b) on face f
        const int fixed = (f & 1) ? 7 : 0;
        if (f < 2) return fixed + 8 * a + 64 * b;
        if (f < 4) return a + 8 * fixed + 64 * b;
        return a + 8 * b + 64 * fixed;
    };
    for (int64_t el = 0; el < nel; el++)
        for (int f = 0; f < 6; f++) {
            FaceLink L{-1, 0, 0};
            const int32_t base_node = (int32_t)(el * GS_ST_N3);
            int npair = 0;
            for (int b = 1; b <= 6; b++)
                for (int a = 1; a <= 6; a++) npair += partner[base_node + node_of(f, a, b)] >= 0;
            bool ok = npair == 36;
            if (ok) {
                const int64_t p11 = partner[base_node + node_of(f, 1, 1)], p21 = partner[base_node + node_of(f, 2, 1)],
                              p12 = partner[base_node + node_of(f, 1, 2)];
                const int64_t sa = p21 - p11, sb = p12 - p11, base = p11 - sa - sb;
                ok = sa >= -32768 && sa <= 32767 && sb >= -32768 && sb <= 32767 && base >= 0 && base <= 2147483647;
                for (int b = 1; b <= 6 && ok; b++)
                    for (int a = 1; a <= 6 && ok; a++) ok = partner[base_node + node_of(f, a, b)] == base + a * sa + b * sb;
                if (ok) L = FaceLink{(int32_t)base, (int16_t)sa, (int16_t)sb};
            }
            if (!ok && npair > 0)
                for (int b = 1; b <= 6; b++)
                    for (int a = 1; a <= 6; a++) {
                        const int32_t nd = base_node + node_of(f, a, b);
                        if (partner[nd] >= 0) gclass[gof[nd]] = GS_S;
                    }
            ftab[(size_t)el * 6 + f] = L;
        }
    // a demoted group takes the face(s) of its other member down with it: sweep until every linked face has all 36 of its
    // groups in P (one sweep on a conforming mesh)
    for (bool changed = true; changed;) {
        changed = false;
        for (int64_t el = 0; el < nel; el++)
            for (int f = 0; f < 6; f++) {
                FaceLink &L = ftab[(size_t)el * 6 + f];
                if (L.base < 0) continue;
                bool ok = true;
                for (int b = 1; b <= 6 && ok; b++)
                    for (int a = 1; a <= 6 && ok; a++) ok = gclass[gof[(int32_t)(el * GS_ST_N3) + node_of(f, a, b)]] == GS_P;
                if (ok) continue;
                for (int b = 1; b <= 6; b++)
                    for (int a = 1; a <= 6; a++) gclass[gof[(int32_t)(el * GS_ST_N3) + node_of(f, a, b)]] = GS_S;
                L = FaceLink{-1, 0, 0};
                changed = true;
            }
    }
    // pass 3: the two CSR subsets and the edge table
    std::vector<int32_t> offE(1, 0), idxE, offS(1, 0), idxS, etab((size_t)nel * GS_ST_EDGE_SLOTS, -1);
    for (int64_t g = 0; g < ng; g++) {
        const int b = off[g], e = off[g + 1];
        if (gclass[g] == GS_E) {
            const int32_t slot = (int32_t)offE.size() - 1;
            for (int q = b; q < e; q++) {
                const int nd = idx[q], l = nd & (GS_ST_N3 - 1);
                etab[(size_t)(nd >> 9) * GS_ST_EDGE_SLOTS + gs_st_slot(l & 7, (l >> 3) & 7, l >> 6)] = slot;
                idxE.push_back(nd);
            }
            offE.push_back((int32_t)idxE.size());
        } else if (gclass[g] == GS_S) {
            idxS.insert(idxS.end(), idx.begin() + b, idx.begin() + e);
            offS.push_back((int32_t)idxS.size());
        }
    }
    h.ngroupsE = (int64_t)offE.size() - 1, h.ngroupsS = (int64_t)offS.size() - 1;
    if (idxE.empty()) idxE.push_back(0);
    if (idxS.empty()) idxS.push_back(0);
    h.ftab.upload(ftab.data(), ftab.size(), s);
    h.etab.upload(etab.data(), etab.size(), s);
    h.goffE.upload(offE.data(), offE.size(), s), h.gidxE.upload(idxE.data(), idxE.size(), s);
    h.goffS.upload(offS.data(), offS.size(), s), h.gidxS.upload(idxS.data(), idxS.size(), s);
    h.gval.alloc((size_t)(h.ngroupsE > 0 ? h.ngroupsE : 1));
    NEKB_CUDA(cudaStreamSynchronize(s));
    h.struct_state = 1;
    return true;
}

// gval[g] = combined value of group g of the E subset (members in ascending order: the bits gs_local_kernel would write)
__global__ void __launch_bounds__(256)
    gs_gval_kernel(double *__restrict__ gval, const double *__restrict__ u, const int32_t *__restrict__ goff,
                   const int32_t *__restrict__ gidx, int ngroups)
{
    for (int gI = blockIdx.x * blockDim.x + threadIdx.x; gI < ngroups; gI += gridDim.x * blockDim.x) {
        const int b = goff[gI], e = goff[gI + 1];
        double v = u[gidx[b]];
        for (int q = b + 1; q < e; q++) v += u[gidx[q]];
        gval[gI] = v;
    }
}

// u <- gs_op(u) [* mask]
inline void gs_op(int handle, double *u, int op, const double *mask)
{
    Ctx &c = ctx();
    GsMap &h = gs_get(handle);
    gs_local(h, u, op);
    if (h.nshared > 0 || c.nranks > 1) gs_remote_exchange(h, u, op);
    if (mask != nullptr && h.n > 0) {
        col2_kernel<<<blocks_for(h.n), 256, 0, c.stream>>>(u, mask, h.n);
        NEKB_LAUNCHED();
    }
}

}  // namespace nekb
