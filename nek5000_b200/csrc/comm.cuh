// comm.cuh -- multi-rank plumbing: NCCL (loaded at run time), host collectives for setup, the
// remote half of gs_op, and the CG all-reduces.
//
// Replaces gop -> mpi_allreduce (core/comm_mpi.f:216-259) and gslib's pairwise exchange
// (gs_op on ids shared between ranks).  One process per GPU (Nek's rank model); NCCL moves the data
// over NVLink 5 / NVSwitch.  libnccl is dlopen'ed so the library loads (and its host-only entry
// points work) on machines without NCCL or without a GPU.
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <map>

#include "gs.cuh"

namespace nekb {

// ---- minimal NCCL declarations (ABI-stable subset of nccl.h 2.x) -----------------------------------------
typedef struct ncclComm *nccl_comm_t;
struct nccl_unique_id {
    char internal[128];
};
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };
enum { NCCL_CHAR = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

inline NcclApi &nccl()
{
    static NcclApi a;
    if (a.lib) return a;
    // By soname first: when the host program (e.g. torch) already loaded its NCCL, this returns that copy.
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    NEKB_REQUIRE(a.lib != nullptr, "NCCL (libnccl.so.2) could not be loaded");
#define NEKB_SYM(field, name)                                            \
    *(void **)(&a.field) = dlsym(a.lib, name);                           \
    NEKB_REQUIRE(a.field != nullptr, std::string("NCCL symbol missing: ") + name)
    NEKB_SYM(GetUniqueId, "ncclGetUniqueId");
    NEKB_SYM(CommInitRank, "ncclCommInitRank");
    NEKB_SYM(CommDestroy, "ncclCommDestroy");
    NEKB_SYM(AllReduce, "ncclAllReduce");
    NEKB_SYM(AllGather, "ncclAllGather");
    NEKB_SYM(Send, "ncclSend");
    NEKB_SYM(Recv, "ncclRecv");
    NEKB_SYM(GroupStart, "ncclGroupStart");
    NEKB_SYM(GroupEnd, "ncclGroupEnd");
    NEKB_SYM(GetErrorString, "ncclGetErrorString");
#undef NEKB_SYM
    return a;
}

#define NEKB_NCCL(call)                                                                      \
    do {                                                                                     \
        int r_ = (call);                                                                     \
        if (r_ != 0) ::nekb::fail(__FILE__, __LINE__, std::string("NCCL: ") + nccl().GetErrorString(r_)); \
    } while (0)

inline nccl_comm_t comm_handle() { return (nccl_comm_t)ctx().nccl_comm; }

inline void comm_allreduce(double *dev, int count, int op)
{
    Ctx &c = ctx();
    if (c.nranks <= 1) return;
    NEKB_REQUIRE(c.nccl_comm != nullptr, "nranks > 1 but nekb_comm_init has not been called");
    NEKB_NCCL(nccl().AllReduce(dev, dev, (size_t)count, NCCL_FLOAT64, op, comm_handle(), c.stream));
}
inline void comm_allreduce_sum(double *dev, int count) { comm_allreduce(dev, count, NCCL_SUM); }
inline void comm_allreduce_max(double *dev, int count) { comm_allreduce(dev, count, NCCL_MAX); }
inline void comm_allreduce_min(double *dev, int count) { comm_allreduce(dev, count, NCCL_MIN); }

// All ranks' streams are drained (each rank syncs its own stream, then joins a one-word all-reduce).  Needed before memory
// that peers write into (the peer-memory gs exchange: receive areas, consumed-flags) is released.
inline void comm_quiesce()
{
    Ctx &c = ctx();
    if (!c.stream) return;
    cudaStreamSynchronize(c.stream);
    if (c.nranks > 1 && c.nccl_comm != nullptr && c.sc.p != nullptr) {
        NEKB_NCCL(nccl().AllReduce(&c.sc.p->work[3], &c.sc.p->work[3], 1, NCCL_FLOAT64, NCCL_MAX, comm_handle(), c.stream));
        cudaStreamSynchronize(c.stream);
    }
}
// release a gs handle whose exchange memory may still be written by peers
inline void gs_release(GsMap &h)
{
    if (h.p2p) comm_quiesce();
    h = GsMap();
}

// ---- host collectives used by the setup code (numbering, shared-id discovery) ----------------------------
// Callbacks registered with nekb_set_transport win (MPI in a Fortran build, gloo in the CPU tests); otherwise
// the bytes are staged through device memory and moved with NCCL.
inline void host_allgather(const void *send, void *recv, size_t bytes)
{
    Ctx &c = ctx();
    if (c.nranks <= 1) {
        memcpy(recv, send, bytes);
        return;
    }
    if (c.allgather) {
        NEKB_REQUIRE(c.allgather(send, recv, bytes, c.transport_user) == 0, "transport allgather failed");
        return;
    }
    NEKB_REQUIRE(c.nccl_comm != nullptr, "no transport: call nekb_set_transport or nekb_comm_init");
    DevBuf<char> s, r;
    s.alloc(bytes ? bytes : 1);
    r.alloc((bytes ? bytes : 1) * c.nranks);
    NEKB_CUDA(cudaMemcpyAsync(s.p, send, bytes, cudaMemcpyHostToDevice, c.stream));
    NEKB_NCCL(nccl().AllGather(s.p, r.p, bytes, NCCL_CHAR, comm_handle(), c.stream));
    NEKB_CUDA(cudaMemcpyAsync(recv, r.p, bytes * c.nranks, cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
}

// send: concatenation of the per-destination blocks (send_bytes[r] each, rank order); recv likewise.
inline void host_alltoallv(const void *send, const int64_t *send_bytes, void *recv, const int64_t *recv_bytes)
{
    Ctx &c = ctx();
    if (c.nranks <= 1) {
        memcpy(recv, send, (size_t)send_bytes[0]);
        return;
    }
    if (c.alltoallv) {
        NEKB_REQUIRE(c.alltoallv(send, send_bytes, recv, recv_bytes, c.transport_user) == 0,
                     "transport alltoallv failed");
        return;
    }
    NEKB_REQUIRE(c.nccl_comm != nullptr, "no transport: call nekb_set_transport or nekb_comm_init");
    int64_t st = 0, rt = 0;
    for (int r = 0; r < c.nranks; r++) st += send_bytes[r], rt += recv_bytes[r];
    DevBuf<char> s, rv;
    s.alloc(st ? st : 1);
    rv.alloc(rt ? rt : 1);
    NEKB_CUDA(cudaMemcpyAsync(s.p, send, (size_t)st, cudaMemcpyHostToDevice, c.stream));
    NEKB_NCCL(nccl().GroupStart());
    int64_t so = 0, ro = 0;
    for (int r = 0; r < c.nranks; r++) {
        if (send_bytes[r]) NEKB_NCCL(nccl().Send(s.p + so, (size_t)send_bytes[r], NCCL_CHAR, r, comm_handle(), c.stream));
        if (recv_bytes[r]) NEKB_NCCL(nccl().Recv(rv.p + ro, (size_t)recv_bytes[r], NCCL_CHAR, r, comm_handle(), c.stream));
        so += send_bytes[r], ro += recv_bytes[r];
    }
    NEKB_NCCL(nccl().GroupEnd());
    NEKB_CUDA(cudaMemcpyAsync(recv, rv.p, (size_t)rt, cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
}

// Exchange of variable-length records: every rank hands over one vector per destination and gets one
// vector per source.
template <class T>
inline std::vector<std::vector<T>> exchange_records(const std::vector<std::vector<T>> &out)
{
    Ctx &c = ctx();
    const int np = c.nranks;
    std::vector<int64_t> sb(np), rb(np), all((size_t)np * np);
    for (int r = 0; r < np; r++) sb[r] = (int64_t)(out[r].size() * sizeof(T));
    host_allgather(sb.data(), all.data(), sizeof(int64_t) * np);
    for (int r = 0; r < np; r++) rb[r] = all[(size_t)r * np + c.rank];
    int64_t st = 0, rt = 0;
    for (int r = 0; r < np; r++) st += sb[r], rt += rb[r];
    std::vector<char> sbuf((size_t)(st ? st : 1)), rbuf((size_t)(rt ? rt : 1));
    int64_t o = 0;
    for (int r = 0; r < np; r++) {
        if (sb[r]) memcpy(sbuf.data() + o, out[r].data(), (size_t)sb[r]);
        o += sb[r];
    }
    host_alltoallv(sbuf.data(), sb.data(), rbuf.data(), rb.data());
    std::vector<std::vector<T>> in(np);
    o = 0;
    for (int r = 0; r < np; r++) {
        in[r].resize((size_t)rb[r] / sizeof(T));
        if (rb[r]) memcpy(in[r].data(), rbuf.data() + o, (size_t)rb[r]);
        o += rb[r];
    }
    return in;
}

// ---- discovery of ids shared between ranks (host) -------------------------------------------------------------
// gslib finds the ranks sharing an id with a crystal-router pass to the id's "owner" (id mod np); the same
// rendezvous is used here.  Input: this rank's distinct non-zero candidate ids (ascending).  Output: for
// every peer (ascending) the ascending list of ids shared with it -- identical on both sides of a pair, which
// fixes the order of the exchange buffers without further negotiation.
struct SharedIds {
    std::vector<int> peers;
    std::vector<std::vector<int64_t>> ids;  // per peer, ascending
};

inline SharedIds discover_shared_ids(const std::vector<int64_t> &uniq)
{
    Ctx &c = ctx();
    const int np = c.nranks;
    SharedIds out;
    if (np <= 1) return out;
    std::vector<std::vector<int64_t>> to_owner(np);
    for (int64_t id : uniq) to_owner[(int)(id % np)].push_back(id);
    std::vector<std::vector<int64_t>> at_owner = exchange_records(to_owner);
    // (id, source) pairs held by this owner
    struct Rec {
        int64_t id;
        int src;
    };
    std::vector<Rec> recs;
    for (int r = 0; r < np; r++)
        for (int64_t id : at_owner[r]) recs.push_back({id, r});
    std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) { return a.id != b.id ? a.id < b.id : a.src < b.src; });
    // reply: to every source of a multiply-held id, the (id, other rank) pairs
    std::vector<std::vector<int64_t>> reply(np);
    for (size_t s = 0; s < recs.size();) {
        size_t t = s + 1;
        while (t < recs.size() && recs[t].id == recs[s].id) t++;
        if (t - s >= 2)
            for (size_t a = s; a < t; a++)
                for (size_t b = s; b < t; b++)
                    if (a != b) {
                        reply[recs[a].src].push_back(recs[a].id);
                        reply[recs[a].src].push_back((int64_t)recs[b].src);
                    }
        s = t;
    }
    std::vector<std::vector<int64_t>> got = exchange_records(reply);
    std::map<int, std::vector<int64_t>> by_peer;
    for (int r = 0; r < np; r++)
        for (size_t q = 0; q + 1 < got[r].size(); q += 2) by_peer[(int)got[r][q + 1]].push_back(got[r][q]);
    for (auto &kv : by_peer) {
        std::sort(kv.second.begin(), kv.second.end());
        out.peers.push_back(kv.first);
        out.ids.push_back(std::move(kv.second));
    }
    return out;
}

// ---- peer-memory exchange (NVLink 5 / NVSwitch, CUDA IPC) -----------------------------------------------------------
// The exchange of gs_op without a communication library call: every rank exports one device allocation per handle
// (double-buffered receive area + flag words) through CUDA IPC; the PACK kernel stores each value straight into the peer's
// receive area over NVLink and then raises an epoch flag in the peer's memory; the UNPACK kernel waits for the flags of
// its peers, combines, and tells the peers that the buffer of this epoch has been consumed.  A buffer is rewritten two
// exchanges later, and only after the owner's "consumed" flag allows it, so back-to-back gs_ops need no other barrier.
// One-node only (all ranks must see each other's memory); NEKB_GS_P2P=0 or any IPC failure keeps the NCCL send/recv path.
constexpr int GS_P2P_MAXPEERS = 32;
struct GsP2pRecord {
    cudaIpcMemHandle_t handle;
    int64_t nitems;
    int32_t ok, npeers;
    int32_t peers[GS_P2P_MAXPEERS];
    int64_t peer_off[GS_P2P_MAXPEERS + 1];
};

inline void gs_p2p_release(GsMap &h)
{
    h.peer_map.close();
    h.p2p = false;
}

inline void gs_p2p_setup(GsMap &h)
{
    Ctx &c = ctx();
    h.p2p = false;
    if (c.nranks <= 1) return;
    const char *env = getenv("NEKB_GS_P2P");
    const bool want = (!env || atoi(env) != 0) && c.nccl_comm != nullptr;
    const int np = (int)h.peers.size();
    const int64_t nitems = h.peer_off.empty() ? 0 : h.peer_off.back();
    GsP2pRecord mine;
    memset(&mine, 0, sizeof mine);
    mine.nitems = nitems, mine.npeers = np;
    mine.ok = want && np <= GS_P2P_MAXPEERS;
    const size_t bytes = (size_t)2 * nitems * sizeof(double) + (size_t)2 * GS_P2P_MAXPEERS * sizeof(unsigned long long) + 64;
    if (mine.ok) {
        h.xmem.alloc(bytes);
        h.xmem.zero(c.stream);
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        if (cudaIpcGetMemHandle(&mine.handle, h.xmem.p) != cudaSuccess) {
            cudaGetLastError();
            mine.ok = 0;
        }
        for (int p = 0; p < np; p++) mine.peers[p] = h.peers[p];
        for (int p = 0; p <= np; p++) mine.peer_off[p] = h.peer_off[p];
    }
    std::vector<GsP2pRecord> all((size_t)c.nranks);
    host_allgather(&mine, all.data(), sizeof mine);
    bool ok = true;
    for (const GsP2pRecord &r : all) ok = ok && r.ok;   // the same decision on every rank
    if (!ok) {          // the same verdict on every rank (all see the same records): nobody goes on to the second collective
        h.xmem.release();
        return;
    }
    // A rank without peers (np == 0) has nothing to open, but it MUST still join the verdict all-gather below: the ranks
    // that do have peers call it, and a collective that some ranks skip pairs with whatever those ranks call next.
    std::vector<double *> precv(np);
    std::vector<int64_t> pstride(np), myoff(np);
    std::vector<unsigned long long *> parr(np), pdone(np);
    h.peer_map.close();
    h.peer_map.v.assign(np, nullptr);
    for (int p = 0; p < np && ok; p++) {
        const GsP2pRecord &r = all[h.peers[p]];
        int slot = -1;
        for (int q = 0; q < r.npeers; q++)
            if (r.peers[q] == c.rank) slot = q;
        const int64_t cnt = h.peer_off[p + 1] - h.peer_off[p];
        if (slot < 0 || r.peer_off[slot + 1] - r.peer_off[slot] != cnt) {
            ok = false;
            break;
        }
        void *base = nullptr;
        if (cudaIpcOpenMemHandle(&base, r.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        h.peer_map.v[p] = base;
        unsigned char *b = static_cast<unsigned char *>(base);
        unsigned long long *flags = reinterpret_cast<unsigned long long *>(b + (size_t)2 * r.nitems * sizeof(double));
        precv[p] = reinterpret_cast<double *>(b) + r.peer_off[slot];
        pstride[p] = r.nitems;
        myoff[p] = h.peer_off[p];
        parr[p] = flags + slot;
        pdone[p] = flags + GS_P2P_MAXPEERS + slot;
    }
    // every rank must reach the same verdict, or some would wait for flags that never come
    int32_t v = ok ? 1 : 0;
    std::vector<int32_t> vs((size_t)c.nranks);
    host_allgather(&v, vs.data(), sizeof v);
    for (int32_t q : vs) ok = ok && q;
    if (!ok) {
        gs_p2p_release(h);
        h.xmem.release();
        return;
    }
    if (np == 0) {      // nothing to exchange, but in step with the others
        h.p2p = true;
        return;
    }
    std::vector<unsigned char> ip((size_t)nitems);
    for (int p = 0; p < np; p++)
        for (int64_t q = h.peer_off[p]; q < h.peer_off[p + 1]; q++) ip[q] = (unsigned char)p;
    cudaStream_t st = c.stream;
    h.item_peer.upload(ip.data(), ip.size(), st);
    h.d_peer_recv.upload(precv.data(), precv.size(), st);
    h.d_peer_stride.upload(pstride.data(), pstride.size(), st);
    h.d_my_off.upload(myoff.data(), myoff.size(), st);
    h.d_peer_arrived.upload(parr.data(), parr.size(), st);
    h.d_peer_done.upload(pdone.data(), pdone.size(), st);
    NEKB_CUDA(cudaStreamSynchronize(st));
    h.epoch = 0;
    h.p2p = true;
}

__device__ __forceinline__ void gs_spin_until(const volatile unsigned long long *flag, unsigned long long want)
{
    const long long t0 = clock64();
    while (*flag < want) {
        if (clock64() - t0 > 40000000000LL) {  // ~20 s: a peer never arrived -- fail loudly instead of hanging the GPU
            printf("nekb200: gs peer-memory exchange timed out (flag %llu < %llu)\n", (unsigned long long)*flag, want);
            __trap();
        }
    }
}

// pack + send: u -> the peers' receive areas (remote stores over NVLink), then the epoch flag
__global__ void __launch_bounds__(256)
    gs_pack_p2p_kernel(const double *__restrict__ u, const int32_t *__restrict__ item_sid, const int32_t *__restrict__ rep,
                       const unsigned char *__restrict__ item_peer, double *const *__restrict__ peer_recv, const int64_t *__restrict__ peer_stride,
                       const int64_t *__restrict__ my_off, unsigned long long *const *__restrict__ peer_arrived,
                       const unsigned long long *done_flags, unsigned *ticket, int nitems, int npeers, unsigned long long epoch)
{
    __shared__ double *s_dst[GS_P2P_MAXPEERS];
    __shared__ int s_last;
    if (threadIdx.x < npeers) {
        // the buffer of this parity was last used at epoch - 2: wait until the peer has consumed it
        if (epoch > 2) gs_spin_until(done_flags + threadIdx.x, epoch - 2);
        s_dst[threadIdx.x] = peer_recv[threadIdx.x] + (epoch & 1ull) * peer_stride[threadIdx.x] - my_off[threadIdx.x];
    }
    __syncthreads();
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nitems; q += gridDim.x * blockDim.x)
        s_dst[item_peer[q]][q] = u[rep[item_sid[q]]];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicInc(ticket, gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < npeers) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(peer_arrived[threadIdx.x]) = epoch;
    }
}

template <int OP>
__global__ void __launch_bounds__(256)
    gs_unpack_p2p_kernel(double *__restrict__ u, const double *recvbuf, const int32_t *__restrict__ s_off, const int32_t *__restrict__ s_items,
                         const int32_t *__restrict__ s_nbelow, const int32_t *__restrict__ rep, const int32_t *__restrict__ x_goff,
                         const int32_t *__restrict__ x_gidx, int nslots, const unsigned long long *arrived_flags,
                         unsigned long long *const *__restrict__ peer_done, unsigned *ticket, int npeers, unsigned long long epoch)
{
    __shared__ int s_last;
    if (threadIdx.x < npeers) gs_spin_until(arrived_flags + threadIdx.x, epoch);
    __syncthreads();
    __threadfence_system();
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
        const int b = s_off[s], e = s_off[s + 1], nb = s_nbelow[s];
        const double mine = u[rep[s]];
        double v;
        if (nb == 0) {
            v = mine;
            for (int q = b; q < e; q++) v = gs_combine<OP>(v, __ldcg(recvbuf + s_items[q]));
        } else {
            v = __ldcg(recvbuf + s_items[b]);
            for (int q = b + 1; q < b + nb; q++) v = gs_combine<OP>(v, __ldcg(recvbuf + s_items[q]));
            v = gs_combine<OP>(v, mine);
            for (int q = b + nb; q < e; q++) v = gs_combine<OP>(v, __ldcg(recvbuf + s_items[q]));
        }
        for (int q = x_goff[s]; q < x_goff[s + 1]; q++) u[x_gidx[q]] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < npeers) {   // every block has read its part of the buffer: the peers may reuse it
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(peer_done[threadIdx.x]) = epoch;
    }
}

// ---- remote part of a gs handle ----------------------------------------------------------------------------------
// cand_idx: local indices whose ids may live on other ranks (nullptr = every non-zero id).  All local copies of
// a shared id must be among the candidates.
inline void gs_build_remote(GsMap &h, const int64_t *id_host, int64_t n, const int32_t *cand_idx, int64_t ncand)
{
    Ctx &c = ctx();
    h.nshared = 0;
    h.peers.clear();
    h.peer_off.assign(1, 0);
    if (c.nranks <= 1) return;
    struct IdIdx {
        int64_t id;
        int32_t idx;
    };
    std::vector<IdIdx> pairs;
    if (cand_idx) {
        pairs.reserve((size_t)ncand);
        for (int64_t q = 0; q < ncand; q++)
            if (id_host[cand_idx[q]] != 0) pairs.push_back({id_host[cand_idx[q]], cand_idx[q]});
    } else {
        for (int64_t q = 0; q < n; q++)
            if (id_host[q] != 0) pairs.push_back({id_host[q], (int32_t)q});
    }
    std::sort(pairs.begin(), pairs.end(), [](const IdIdx &a, const IdIdx &b) { return a.id != b.id ? a.id < b.id : a.idx < b.idx; });
    std::vector<int64_t> uniq;
    std::vector<int64_t> ustart;  // start of each distinct id in `pairs`
    for (size_t s = 0; s < pairs.size(); s++)
        if (s == 0 || pairs[s].id != pairs[s - 1].id) {
            uniq.push_back(pairs[s].id);
            ustart.push_back((int64_t)s);
        }
    ustart.push_back((int64_t)pairs.size());
    SharedIds sh = discover_shared_ids(uniq);

    // slots: distinct shared ids, ascending
    std::vector<int64_t> sids;
    for (auto &v : sh.ids) sids.insert(sids.end(), v.begin(), v.end());
    std::sort(sids.begin(), sids.end());
    sids.erase(std::unique(sids.begin(), sids.end()), sids.end());
    const int64_t ns = (int64_t)sids.size();
    h.nshared = ns;
    h.peers = sh.peers;
    if (ns == 0) {
        gs_p2p_setup(h);  // collective: ranks without shared ids still take part in the handle exchange
        return;
    }
    NEKB_REQUIRE(ns < 2147483647, "gs_setup: too many shared ids");

    std::vector<int32_t> x_goff(ns + 1), x_gidx, rep(ns);
    for (int64_t s = 0; s < ns; s++) {
        const size_t u = std::lower_bound(uniq.begin(), uniq.end(), sids[s]) - uniq.begin();
        x_goff[s] = (int32_t)x_gidx.size();
        for (int64_t q = ustart[u]; q < ustart[u + 1]; q++) x_gidx.push_back(pairs[q].idx);
        rep[s] = pairs[ustart[u]].idx;
    }
    x_goff[ns] = (int32_t)x_gidx.size();

    // exchange items: per peer, per shared id (ascending) -> slot
    std::vector<int32_t> item_sid;
    for (size_t p = 0; p < sh.peers.size(); p++) {
        for (int64_t id : sh.ids[p])
            item_sid.push_back((int32_t)(std::lower_bound(sids.begin(), sids.end(), id) - sids.begin()));
        h.peer_off.push_back((int64_t)item_sid.size());
    }
    const int64_t nitems = (int64_t)item_sid.size();
    // per slot: the items (recv positions) contributing to it, in ascending peer order, and how many of those
    // peers rank below this rank (self is folded in at that position => same summation order on every rank)
    std::vector<int32_t> s_off(ns + 1, 0), s_items(nitems), s_nbelow(ns, 0);
    for (int64_t q = 0; q < nitems; q++) s_off[item_sid[q] + 1]++;
    for (int64_t s = 0; s < ns; s++) s_off[s + 1] += s_off[s];
    {
        std::vector<int32_t> fill(s_off.begin(), s_off.end() - 1);
        for (size_t p = 0; p < sh.peers.size(); p++)
            for (int64_t q = h.peer_off[p]; q < h.peer_off[p + 1]; q++) {
                s_items[fill[item_sid[q]]++] = (int32_t)q;
                if (sh.peers[p] < c.rank) s_nbelow[item_sid[q]]++;
            }
    }
    cudaStream_t st = c.stream;
    h.x_goff.upload(x_goff.data(), x_goff.size(), st);
    h.x_gidx.upload(x_gidx.data(), x_gidx.size(), st);
    h.x_rep.upload(rep.data(), rep.size(), st);
    h.x_item_sid.upload(item_sid.data(), item_sid.size(), st);
    h.x_soff.upload(s_off.data(), s_off.size(), st);
    h.x_sitems.upload(s_items.data(), s_items.size(), st);
    h.x_nbelow.upload(s_nbelow.data(), s_nbelow.size(), st);
    h.sendbuf.alloc(nitems);
    h.recvbuf.alloc(nitems);
    h.nx_members = (int64_t)x_gidx.size();
    NEKB_CUDA(cudaStreamSynchronize(st));
    gs_p2p_setup(h);
}

__global__ void __launch_bounds__(256)
    gs_pack_kernel(double *__restrict__ sendbuf, const double *__restrict__ u, const int32_t *__restrict__ item_sid,
                   const int32_t *__restrict__ rep, int nitems)
{
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nitems; q += gridDim.x * blockDim.x)
        sendbuf[q] = u[rep[item_sid[q]]];
}

template <int OP>
__global__ void __launch_bounds__(256)
    gs_unpack_kernel(double *__restrict__ u, const double *__restrict__ recvbuf, const int32_t *__restrict__ s_off,
                     const int32_t *__restrict__ s_items, const int32_t *__restrict__ s_nbelow,
                     const int32_t *__restrict__ rep, const int32_t *__restrict__ x_goff,
                     const int32_t *__restrict__ x_gidx, int nslots)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
        const int b = s_off[s], e = s_off[s + 1], nb = s_nbelow[s];
        const double mine = u[rep[s]];
        double v;
        if (nb == 0) {
            v = mine;
            for (int q = b; q < e; q++) v = gs_combine<OP>(v, recvbuf[s_items[q]]);
        } else {
            v = recvbuf[s_items[b]];
            for (int q = b + 1; q < b + nb; q++) v = gs_combine<OP>(v, recvbuf[s_items[q]]);
            v = gs_combine<OP>(v, mine);
            for (int q = b + nb; q < e; q++) v = gs_combine<OP>(v, recvbuf[s_items[q]]);
        }
        for (int q = x_goff[s]; q < x_goff[s + 1]; q++) u[x_gidx[q]] = v;
    }
}

inline void gs_remote_exchange(GsMap &h, double *u, int op)
{
    Ctx &c = ctx();
    if (c.nranks <= 1 || h.nshared == 0) return;
    NEKB_REQUIRE(c.nccl_comm != nullptr, "gs_op: ids are shared between ranks but nekb_comm_init was not called");
    const int nitems = (int)h.peer_off.back();
    if (h.p2p) {
        const int npeers = (int)h.peers.size();
        const unsigned long long epoch = ++h.epoch;
        unsigned char *base = h.xmem.p;
        const double *recv = reinterpret_cast<const double *>(base) + (epoch & 1ull) * (size_t)nitems;
        unsigned long long *flags = reinterpret_cast<unsigned long long *>(base + (size_t)2 * nitems * sizeof(double));
        unsigned *tickets = reinterpret_cast<unsigned *>(flags + 2 * GS_P2P_MAXPEERS);
        int gp = blocks_for(nitems), gu = blocks_for(h.nshared);
        if (gp > c.num_sms * 4) gp = c.num_sms * 4;   // all blocks are co-resident: the flag protocol never waits on an unscheduled block
        if (gu > c.num_sms * 4) gu = c.num_sms * 4;
        gs_pack_p2p_kernel<<<gp, 256, 0, c.stream>>>(u, h.x_item_sid.p, h.x_rep.p, h.item_peer.p, h.d_peer_recv.p, h.d_peer_stride.p,
                                                     h.d_my_off.p, h.d_peer_arrived.p, flags + GS_P2P_MAXPEERS, tickets, nitems, npeers, epoch);
        NEKB_LAUNCHED();
        const int ns2 = (int)h.nshared;
#define NEKB_UNPACK_P2P(OPV)                                                                                                   \
    gs_unpack_p2p_kernel<OPV><<<gu, 256, 0, c.stream>>>(u, recv, h.x_soff.p, h.x_sitems.p, h.x_nbelow.p, h.x_rep.p, h.x_goff.p,   \
                                                        h.x_gidx.p, ns2, flags, h.d_peer_done.p, tickets + 1, npeers, epoch)
        switch (op) {
            case 1: NEKB_UNPACK_P2P(1); break;
            case 2: NEKB_UNPACK_P2P(2); break;
            case 3: NEKB_UNPACK_P2P(3); break;
            default: NEKB_UNPACK_P2P(4); break;
        }
#undef NEKB_UNPACK_P2P
        NEKB_LAUNCHED();
        return;
    }
    gs_pack_kernel<<<blocks_for(nitems), 256, 0, c.stream>>>(h.sendbuf.p, u, h.x_item_sid.p, h.x_rep.p, nitems);
    NEKB_LAUNCHED();
    NEKB_NCCL(nccl().GroupStart());
    for (size_t p = 0; p < h.peers.size(); p++) {
        const int64_t o = h.peer_off[p], cnt = h.peer_off[p + 1] - o;
        NEKB_NCCL(nccl().Send(h.sendbuf.p + o, (size_t)cnt, NCCL_FLOAT64, h.peers[p], comm_handle(), c.stream));
        NEKB_NCCL(nccl().Recv(h.recvbuf.p + o, (size_t)cnt, NCCL_FLOAT64, h.peers[p], comm_handle(), c.stream));
    }
    NEKB_NCCL(nccl().GroupEnd());
    const int ns = (int)h.nshared;
    const int grid = blocks_for(ns);
#define NEKB_UNPACK(OPV)                                                                                         \
    gs_unpack_kernel<OPV><<<grid, 256, 0, c.stream>>>(u, h.recvbuf.p, h.x_soff.p, h.x_sitems.p, h.x_nbelow.p,   \
                                                      h.x_rep.p, h.x_goff.p, h.x_gidx.p, ns)
    switch (op) {
        case 1: NEKB_UNPACK(1); break;
        case 2: NEKB_UNPACK(2); break;
        case 3: NEKB_UNPACK(3); break;
        default: NEKB_UNPACK(4); break;
    }
#undef NEKB_UNPACK
    NEKB_LAUNCHED();
}

}  // namespace nekb
