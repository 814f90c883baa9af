// ctx.cuh -- process-wide state of libnekb200 (one GPU per process, Nek's rank model).
#pragma once
#include "common.cuh"

namespace nekb {

constexpr int MAX_NX = 16;

// Derivative matrix, row-major: c_D[a * nx + b] = D(a,b) = dxm1(a,b) (core/DXYZ:4).
__constant__ double c_D[MAX_NX * MAX_NX];
__constant__ double c_z[MAX_NX];   // GLL points  (uploaded by setup.cuh upload_gll_constants / ax.cuh ax_affine_ensure)
__constant__ double c_w[MAX_NX];   // GLL weights

// Device-side scalars of the CG drivers (one cache line region, zero-initialised).
struct CgScalars {
    int it;                // iteration counter advanced by the last block of the closing kernel
    int done;              // convergence flag (stock cggo path)
    int niter;             // iterations performed at convergence
    int pad0;
    unsigned counter[4];   // grid_reduce tickets (one per kernel kind)
    double rtz1, rtz2;     // cggo: (z,r) of this / previous iteration
    double rho;            // cggo: (w,p)
    double rbn2, rbn0, tol;
    double alpha;          // cggos (fused path): step length of the last completed iteration
    double work[4];
};

// CUDA IPC mappings of peer allocations, closed when the owning handle goes away
struct IpcMaps {
    std::vector<void *> v;
    IpcMaps() = default;
    IpcMaps(const IpcMaps &) = delete;
    IpcMaps &operator=(const IpcMaps &) = delete;
    IpcMaps(IpcMaps &&o) noexcept : v(std::move(o.v)) { o.v.clear(); }
    IpcMaps &operator=(IpcMaps &&o) noexcept
    {
        if (this != &o) {
            close();
            v = std::move(o.v);
            o.v.clear();
        }
        return *this;
    }
    ~IpcMaps() { close(); }
    void close()
    {
        for (void *q : v)
            if (q) cudaIpcCloseMemHandle(q);
        v.clear();
    }
};

// Node classes of an 8^3 tile by the number of local indices on the tile surface: 0 interior, 1 face-interior, >= 2 edge /
// corner.  Slots of the edge / corner nodes in the per-element table: x-edges 0..23 (edge (j,k in {0,7}) * 6 + i-1), y-edges
// 24..47, z-edges 48..71, corners 72..79.
constexpr int GS_ST_NX = 8, GS_ST_N3 = 512, GS_ST_EDGE_SLOTS = 80;
__host__ __device__ __forceinline__ int gs_st_nb(int i, int j, int k)
{
    return (i == 0 || i == 7) + (j == 0 || j == 7) + (k == 0 || k == 7);
}
__host__ __device__ __forceinline__ int gs_st_slot(int i, int j, int k)   // nb >= 2 only
{
    const int bi = (i == 0 || i == 7), bj = (j == 0 || j == 7), bk = (k == 0 || k == 7);
    const int hi = i == 7, hj = j == 7, hk = k == 7;
    if (bi + bj + bk == 3) return 72 + hi + 2 * hj + 4 * hk;
    if (!bi) return (hj + 2 * hk) * 6 + (i - 1);
    if (!bj) return 24 + (hi + 2 * hk) * 6 + (j - 1);
    return 48 + (hi + 2 * hj) * 6 + (k - 1);
}
// nb == 1 only: face (0..5 = -x,+x,-y,+y,-z,+z) and the in-face coordinates (lower axis first)
__host__ __device__ __forceinline__ int gs_st_face(int i, int j, int k, int &a, int &b)
{
    if (i == 0 || i == 7) {
        a = j, b = k;
        return i == 7;
    }
    if (j == 0 || j == 7) {
        a = i, b = k;
        return 2 + (j == 7);
    }
    a = i, b = j;
    return 4 + (k == 7);
}


// partner(a, b) = base + a * sa + b * sb for the node with in-face coordinates (a, b) (the two free local indices, lower axis first)
struct FaceLink {
    int32_t base;
    int16_t sa, sb;
};

// Local gather-scatter map (gslib gs_setup result restated for the device).
struct GsMap {
    bool used = false;
    int64_t n = 0;                 // vector length the handle was set up for
    int64_t ngroups = 0;           // ids carried by >= 2 local entries
    int64_t nmembers = 0;          // sum of group sizes
    DevBuf<int32_t> goff;          // [ngroups+1] CSR offsets, groups ordered by first member
    DevBuf<int32_t> gidx;          // [nmembers] member indices, ascending inside a group
    DevBuf<int32_t> link;          // [n] node -> its group, for gather-style consumers (gs.cuh gs_ensure_link): -1 not in a
                                   // group, >= 0 the other member of a pair, <= -2 group -2-g (three or more members)
    int link_mode = 0;             // 0 not built, 1 every group through `link`, 2 pairs through `link` + goff3/gidx3
    DevBuf<int32_t> goff3, gidx3;  // CSR of the groups with three or more members (edges, corners), link_mode 2
    int64_t ngroups3 = 0;
    // ---- structured gather view of the same map for the fused cggos update (gs.cuh gs_ensure_struct, lx1 = 8) ----------
    // Face-interior nodes of a conforming hex mesh pair face to face by an affine index map, so a face needs 8 bytes
    // (base, two strides) instead of a 4-byte partner per node; edge / corner nodes read the assembled value of their group
    // from `gval` through an 80-entry table per element; everything else (groups with remote members, non-conforming
    // leftovers) is assembled in place by the stock kernel on the (goffS, gidxS) subset.
    int struct_state = 0;          // 0 not built, 1 built, -1 not available for this handle (vector is not nel * 8^3 nodes)
    DevBuf<FaceLink> ftab;         // [nel][6]  faces -x,+x,-y,+y,-z,+z; base < 0: no gathered partner on that face
    DevBuf<int32_t> etab;          // [nel][80] edge / corner node -> slot in gval, or -1 (own value is final)
    DevBuf<int32_t> goffE, gidxE;  // CSR of the groups assembled into gval (all members are edge / corner nodes, no remote member)
    DevBuf<int32_t> goffS, gidxS;  // CSR of the groups assembled in place
    DevBuf<double> gval;           // [ngroupsE]
    int64_t ngroupsE = 0, ngroupsS = 0;
    // ---- remote part (np > 1): ids shared with other ranks --------------------------------------
    int64_t nshared = 0;           // local unique ids that also live on another rank
    std::vector<int> peers;        // neighbour ranks, ascending
    std::vector<int64_t> peer_off; // [npeers+1] offsets into the exchange lists
    DevBuf<int32_t> x_goff;        // [nshared+1] scatter CSR: local members of every shared id
    DevBuf<int32_t> x_gidx;
    DevBuf<int32_t> x_rep;         // [nshared] one local member per shared id (holds the local sum)
    DevBuf<int32_t> x_item_sid;    // [nitems] shared-id slot of every exchange item (peer-major, id ascending)
    DevBuf<int32_t> x_soff;        // [nshared+1] CSR slot -> contributing recv items, ascending peer
    DevBuf<int32_t> x_sitems;
    DevBuf<int32_t> x_nbelow;      // [nshared] how many of those peers have a lower rank than this one
    DevBuf<double> sendbuf, recvbuf;
    int64_t nx_members = 0;
    // ---- peer-memory exchange over NVLink (CUDA IPC; comm.cuh gs_p2p_*) -------------------------------------------
    bool p2p = false;
    uint64_t epoch = 0;                // exchanges done with this handle (identical on every rank)
    DevBuf<unsigned char> xmem;        // [2][nitems] doubles (double-buffered receive) | arrived[npeers] | done[npeers] | tickets
    IpcMaps peer_map;                  // IPC mappings of the peers' xmem (closed with the handle)
    DevBuf<unsigned char> item_peer;   // [nitems] peer index of every exchange item
    DevBuf<double *> d_peer_recv;      // [npeers] peer's receive area for my segment (buffer 0)
    DevBuf<int64_t> d_peer_stride;     // [npeers] peer's nitems (distance between its two buffers)
    DevBuf<int64_t> d_my_off;          // [npeers] start of the segment in my own item list
    DevBuf<unsigned long long *> d_peer_arrived, d_peer_done;   // [npeers] my slots in the peer's flag arrays
};

struct Ctx {
    bool inited = false;
    int device = 0;
    int nx = 0, nxyz = 0;
    int nelv = 0, nelt = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string last_error;
    void (*exit_handler)(void) = nullptr;

    // registered state ----------------------------------------------------------------------------
    std::vector<double> D_host;    // row-major D(a,b)
    bool have_D = false;
    std::vector<double> z_host, w_host;  // GLL points / weights (zgm1, wxm1)
    bool have_gll = false;
    DevBuf<double> g;              // [nelt][6][nxyz], order rr,rs,rt,ss,st,tt
    DevBuf<double> bm1;            // [nelt][nxyz]
    DevBuf<double> v1mask;         // bp5 mask
    std::vector<int> ifdfrm;       // empty = all deformed
    bool have_geom = false;
    // Elements whose six factors are a per-element constant times w_i w_j w_k (affine elements: every genbox brick, stretched
    // or not): the BP5 operator kernel then reads 6 doubles per ELEMENT instead of 6 per NODE (ax.cuh ax_cg_affine_kernel).
    int geom_gen = 0;              // bumped whenever `g` changes
    int affine_gen = -1;           // geom_gen the verdict below belongs to
    bool affine = false;           // every element passed the test
    double affine_maxdev = 0.0;    // largest relative deviation of a registered factor from constant * w3 (last check)
    DevBuf<double> gc;             // [nelt][8]: rr,rs,rt,ss,st,tt constants (+2 pad: 64-byte records for bulk copies)
    int ifield = 1;
    int gsh_fld[32];
    int istep = 0;
    double volvm1 = 0.0, voltm1 = 0.0;
    int niterhm = 0;
    // residual history of the most recent cggo / hmh_gmres / hmh_flex_cg solve, whoever called it (hmholtz_, hsolve_, ...):
    // what the reference only prints (hmholtz.f:770-773, gmres.f:496-498); read back with nekb_last_history
    std::vector<double> last_hist;
    int last_hist_rows = 0, last_hist_cols = 0;
    double param[201] = {0};       // INPUT param(1:200) entries the path reads (18, 21, 22); 1-based
    double restol[32] = {0};       // TSTEP restol(0:ldimt1): per-field residual tolerance that overrules cggo's tin (hmholtz.f:676)
    DevBuf<double> binvm1, bintm1; // MASS binvm1 / bintm1 for hmholtz
    DevBuf<double> vmask[3], vmult; // SOLN v1mask,v2mask,v3mask and vmult for ophinv
    int niter3[3] = {0, 0, 0};     // niterhm of the three component solves of the last ophinv
    bool ifprojfld[16] = {false};  // INPUT ifprojfld(0:ldimt1): residual projection per field (hsolve)
    int ldimt_proj = 3;            // SIZE ldimt_proj

    // gs handles ---------------------------------------------------------------------------------
    std::vector<GsMap> gs;

    // reduction scratch ---------------------------------------------------------------------------
    DevBuf<double> partials;       // >= 4 * max grid
    DevBuf<CgScalars> sc;
    DevBuf<double> hist;           // device history arrays
    DevBuf<unsigned char> wcode;   // per-node weight/mask codes of the fused cggos path
    DevBuf<int> flags;             // small device flag words
    DevBuf<double> work[8];        // CG work vectors (r, p, w, d, ...)
    DevBuf<double> stage[8];       // device staging of the host-buffer (Fortran-named) entry points

    // multi-rank ----------------------------------------------------------------------------------
    int rank = 0, nranks = 1;
    void *transport_user = nullptr;
    int (*allgather)(const void *, void *, size_t, void *) = nullptr;
    int (*alltoallv)(const void *, const int64_t *, void *, const int64_t *, void *) = nullptr;
    void *nccl_comm = nullptr;
};

// Optional per-kernel timing with CUDA events on the library stream (bench.py's live roofline numbers).
enum ProfCat { PROF_AX = 0, PROF_GS = 1, PROF_UPDATE = 2, PROF_PUPDATE = 3, PROF_NCAT = 4 };
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Span {
        int cat;
        size_t e0, e1;
    };
    std::vector<Span> spans;
    double secs[PROF_NCAT] = {0, 0, 0, 0};
    int64_t count[PROF_NCAT] = {0, 0, 0, 0};
    size_t open_ev[PROF_NCAT] = {0, 0, 0, 0};
};
inline Prof &prof()
{
    static Prof p;
    return p;
}

inline Ctx &ctx()
{
    static Ctx c;
    return c;
}

inline size_t prof_event(cudaStream_t s)
{
    Prof &p = prof();
    if (p.used == p.pool.size()) {
        cudaEvent_t e;
        NEKB_CUDA(cudaEventCreate(&e));
        p.pool.push_back(e);
    }
    NEKB_CUDA(cudaEventRecord(p.pool[p.used], s));
    return p.used++;
}
inline void prof_begin(int cat)
{
    if (prof().on) prof().open_ev[cat] = prof_event(ctx().stream);
}
inline void prof_end(int cat)
{
    Prof &p = prof();
    if (p.on) p.spans.push_back({cat, p.open_ev[cat], prof_event(ctx().stream)});
}
// Call after the stream has been synchronised.
inline void prof_collect()
{
    Prof &p = prof();
    for (const Prof::Span &sp : p.spans) {
        float ms = 0.f;
        NEKB_CUDA(cudaEventElapsedTime(&ms, p.pool[sp.e0], p.pool[sp.e1]));
        p.secs[sp.cat] += 1e-3 * ms;
        p.count[sp.cat]++;
    }
    p.spans.clear();
    p.used = 0;
}

inline void require_init()
{
    NEKB_REQUIRE(ctx().inited, "nekb_init has not been called");
}

// Persistent-grid size: a multiple of the SM count (B200: 148), capped by the work available.
inline int grid_for(int64_t work_items, int ctas_per_sm)
{
    int64_t g = (int64_t)ctx().num_sms * ctas_per_sm;
    if (g > work_items) g = work_items;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace nekb
