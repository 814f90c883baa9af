// readers.cuh -- host-side readers of the reference's binary mesh (.re2) and partition/vertex (.ma2) files, and the
// element -> rank assignment, so that the harness can run the reference's own fixtures (examples/bp5/bp5.{re2,ma2},
// short_tests/ethier, examples/turbChannel) instead of synthetic boxes only.
//
//   .re2  core/reader_re2.f:543-639 (header), :65-158 + :391-471 (mesh records), :160-290 (curved sides), :292-389 +
//         :473-541 (boundary conditions)
//   .ma2  core/map2.f:712-941 (read_map), :943-1026 (assign_gllnid), core/math.f isort / iswapt_ip
//
// Plain host code: file I/O is not a GPU job; what it produces (corner coordinates, vertex ids, element->rank map) is
// what setupds / the geometry kernels consume.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace nekb {

struct Re2Header {
    int version = 0;        // 1, 2, 3
    int64_t nelgt = 0, nelgv = 0;
    int ldim = 0;
    int wdsize = 4;         // bytes per record word: 4 (#v001) or 8 (#v002, #v003)
    bool swap = false;      // endian tag 6.54321 read back byte-reversed
};

inline uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
inline uint64_t bswap64(uint64_t v)
{
    return ((uint64_t)bswap32((uint32_t)v) << 32) | bswap32((uint32_t)(v >> 32));
}

struct FileHandle {
    FILE *f = nullptr;
    explicit FileHandle(const char *path) : f(fopen(path, "rb")) {}
    ~FileHandle()
    {
        if (f) fclose(f);
    }
    void read_at(int64_t off, void *dst, size_t bytes, const char *what)
    {
        NEKB_REQUIRE(fseeko(f, (off_t)off, SEEK_SET) == 0 && fread(dst, 1, bytes, f) == bytes,
                     std::string("short read of ") + what);
    }
};

// if_byte_swap_test (core/byte_mpi.f / prepost): the tag is 6.54321 stored as real*4
inline bool endian_tag_swapped(float test, const char *what)
{
    const float tag = 6.54321f;
    if (fabsf(test - tag) < 1e-5f) return false;
    uint32_t u;
    memcpy(&u, &test, 4);
    u = bswap32(u);
    float t2;
    memcpy(&t2, &u, 4);
    NEKB_REQUIRE(fabsf(t2 - tag) < 1e-5f, std::string(what) + ": endian tag is not 6.54321");
    return true;
}

// core/reader_re2.f:543-639 read_re2_hdr: 80-byte header '#v00x' + nelgt, ldim, nelgv; then the endian tag
inline Re2Header re2_header(FileHandle &fh)
{
    char hdr[81] = {0};
    fh.read_at(0, hdr, 80, ".re2 header");
    Re2Header h;
    NEKB_REQUIRE(!strncmp(hdr, "#v00", 4) && hdr[4] >= '1' && hdr[4] <= '4', ".re2: unknown header version");
    h.version = hdr[4] - '0';
    long long a = 0, c = 0;
    int b = 0;
    if (h.version == 4) {
        // '#v004': list-directed read of version, nelgt, ldimr, nelgv, nBCre2 (core/reader_re2.f:585-586)
        long long nbc4 = 0;
        NEKB_REQUIRE(sscanf(hdr + 5, "%lld %d %lld %lld", &a, &b, &c, &nbc4) >= 3, ".re2: cannot parse the #v004 header");
    } else {
        // format (a5,i9,i3,i9) (core/reader_re2.f:588-590): FIXED columns -- with nelgt >= 10^8 the i3 and i9 fields abut
        // ('  3123456789'), so the fields are cut by position, not by white space
        auto field = [&](int lo, int hi) -> long long {
            char buf[16] = {0};
            memcpy(buf, hdr + lo, (size_t)(hi - lo));
            char *end = nullptr;
            const long long v = strtoll(buf, &end, 10);
            while (end && *end == ' ') end++;
            NEKB_REQUIRE(end != buf && end && *end == 0, ".re2: cannot parse the header (format a5,i9,i3,i9)");
            return v;
        };
        a = field(5, 14), b = (int)field(14, 17), c = field(17, 26);
    }
    h.nelgt = a, h.ldim = b, h.nelgv = c;
    h.wdsize = h.version == 1 ? 4 : 8;   // reader_re2.f:593-596
    NEKB_REQUIRE(h.ldim == 2 || h.ldim == 3, ".re2: ldim must be 2 or 3");
    float test;
    fh.read_at(80, &test, 4, ".re2 endian tag");
    h.swap = endian_tag_swapped(test, ".re2");
    return h;
}

// one record word -> double (core/reader_re2.f buf_to_xyz :391-471: real*4 words for #v001, real*8 otherwise)
inline double re2_word(const unsigned char *p, int wdsize, bool swap)
{
    if (wdsize == 8) {
        uint64_t u;
        memcpy(&u, p, 8);
        if (swap) u = bswap64(u);
        double d;
        memcpy(&d, &u, 8);
        return d;
    }
    uint32_t u;
    memcpy(&u, p, 4);
    if (swap) u = bswap32(u);
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
}

struct Re2Sections {
    int64_t mesh_off = 84, curve_count_off = 0, ncurve = 0, curve_off = 0;
    std::vector<int64_t> bc_count_off, nbc, bc_off;  // one entry per field present in the file
};

// Walks the section table: mesh records, curve count + records, then (count + records) per field until end of file.
inline Re2Sections re2_sections(FileHandle &fh, const Re2Header &h)
{
    Re2Sections s;
    const int nv = 1 << h.ldim;
    const int64_t lrs = 1 + (int64_t)h.ldim * nv;
    fseeko(fh.f, 0, SEEK_END);
    const int64_t fsize = (int64_t)ftello(fh.f);
    int64_t off = 84 + h.nelgt * lrs * h.wdsize;
    auto count_at = [&](int64_t o) -> int64_t {
        unsigned char b[8];
        fh.read_at(o, b, (size_t)h.wdsize, ".re2 record count");
        if (h.wdsize == 8) return (int64_t)re2_word(b, 8, h.swap);
        uint32_t u;
        memcpy(&u, b, 4);
        if (h.swap) u = bswap32(u);
        return (int64_t)(int32_t)u;
    };
    s.curve_count_off = off;
    s.ncurve = count_at(off);
    off += h.wdsize;
    s.curve_off = off;
    off += s.ncurve * 8 * h.wdsize;
    while (off + h.wdsize <= fsize) {
        const int64_t n = count_at(off);
        NEKB_REQUIRE(n >= 0 && off + h.wdsize + n * 8 * h.wdsize <= fsize, ".re2: boundary-condition section overruns the file");
        s.bc_count_off.push_back(off);
        s.nbc.push_back(n);
        s.bc_off.push_back(off + h.wdsize);
        off += h.wdsize + n * 8 * h.wdsize;
    }
    return s;
}

// Elements [e0, e0+nel) in file (= global) order: group id and the 2^ldim corner coordinates in PREPROCESSOR order, as
// xc(8,e), yc(8,e), zc(8,e) of core/INPUT (buf_to_xyz, reader_re2.f:391-471).
inline void re2_read_mesh(FileHandle &fh, const Re2Header &h, int64_t e0, int64_t nel, double *xc, double *yc, double *zc, int *igroup)
{
    NEKB_REQUIRE(e0 >= 0 && nel >= 0 && e0 + nel <= h.nelgt, ".re2: element range outside the file");
    const int nv = 1 << h.ldim;
    const int64_t lrs = 1 + (int64_t)h.ldim * nv, rec = lrs * h.wdsize;
    std::vector<unsigned char> buf((size_t)(nel * rec));
    if (nel) fh.read_at(84 + e0 * rec, buf.data(), buf.size(), ".re2 mesh records");
    for (int64_t e = 0; e < nel; e++) {
        const unsigned char *p = buf.data() + e * rec;
        if (igroup) igroup[e] = (int)re2_word(p, h.wdsize, h.swap);
        for (int v = 0; v < nv; v++) {
            xc[e * nv + v] = re2_word(p + (size_t)(1 + v) * h.wdsize, h.wdsize, h.swap);
            yc[e * nv + v] = re2_word(p + (size_t)(1 + nv + v) * h.wdsize, h.wdsize, h.swap);
            if (h.ldim == 3) zc[e * nv + v] = re2_word(p + (size_t)(1 + 2 * nv + v) * h.wdsize, h.wdsize, h.swap);
        }
    }
}

// Boundary conditions of one field section (readp_re2_bc + buf_to_bc, reader_re2.f:292-389,473-541): record = element,
// face, bl(5), cbl (3 characters in the last word, never byte-swapped).  Outputs are global arrays cbc[6*nelgt][3] (blank
// filled by the caller with 'E  ' semantics of an absent record) and bc[5*6*nelgt].
inline void re2_read_bc(FileHandle &fh, const Re2Header &h, const Re2Sections &s, int section, char *cbc, double *bc)
{
    NEKB_REQUIRE(section >= 0 && section < (int)s.nbc.size(), ".re2: no such boundary-condition section");
    const int64_t n = s.nbc[section], rec = 8 * (int64_t)h.wdsize;
    std::vector<unsigned char> buf((size_t)(n * rec));
    if (n) fh.read_at(s.bc_off[section], buf.data(), buf.size(), ".re2 boundary records");
    for (int64_t r = 0; r < n; r++) {
        const unsigned char *p = buf.data() + r * rec;
        int64_t eg, f;
        if (h.wdsize == 8) {
            eg = (int64_t)re2_word(p, 8, h.swap);
            f = (int64_t)re2_word(p + 8, 8, h.swap);
        } else {
            uint32_t a, b;
            memcpy(&a, p, 4), memcpy(&b, p + 4, 4);
            if (h.swap) a = bswap32(a), b = bswap32(b);
            eg = (int32_t)a, f = (int32_t)b;
        }
        NEKB_REQUIRE(eg >= 1 && eg <= h.nelgt && f >= 1 && f <= 2 * h.ldim, ".re2: bad boundary record");
        const int64_t slot = (eg - 1) * 6 + (f - 1);
        if (bc)
            for (int k = 0; k < 5; k++) bc[slot * 5 + k] = re2_word(p + (size_t)(2 + k) * h.wdsize, h.wdsize, h.swap);
        if (cbc) memcpy(cbc + slot * 3, p + (size_t)7 * h.wdsize, 3);
        // reader_re2.f:533-534: in a 4-byte file with nelgt >= 10^6 the first word of a periodic ('P  ') record is the
        // partner element as an INTEGER (a real*4 cannot hold it exactly): bl(1) = buf(3) as int.  (8-byte files, :522-523, copy
        // the same bits with copyi4, which the double read above reproduces only for the 4-byte case -- they carry the id as a
        // double there.)
        if (bc && h.wdsize == 4 && h.nelgt >= 1000000 && !memcmp(p + (size_t)7 * h.wdsize, "P  ", 3)) {
            uint32_t w;
            memcpy(&w, p + (size_t)2 * h.wdsize, 4);
            if (h.swap) w = bswap32(w);
            bc[slot * 5] = (double)(int32_t)w;
        }
    }
}

// Curved sides (readp_re2_curve + buf_to_curve, reader_re2.f:160-290,473-497): record = element, side (edge 1..12 or face
// 1..6 depending on the curve type), curve(5), ccurve (one character in the last word, never byte-swapped).  Outputs are
// the global arrays curve[5*12*nelgt] and ccurve[12*nelgt] of core/INPUT (slots without a record are left untouched).
inline void re2_read_curves(FileHandle &fh, const Re2Header &h, const Re2Sections &s, char *ccurve, double *curve)
{
    const int64_t n = s.ncurve, rec = 8 * (int64_t)h.wdsize;
    std::vector<unsigned char> buf((size_t)(n * rec));
    if (n) fh.read_at(s.curve_off, buf.data(), buf.size(), ".re2 curved-side records");
    for (int64_t r = 0; r < n; r++) {
        const unsigned char *p = buf.data() + r * rec;
        int64_t eg, f;
        if (h.wdsize == 8) {
            eg = (int64_t)re2_word(p, 8, h.swap);
            f = (int64_t)re2_word(p + 8, 8, h.swap);
        } else {
            uint32_t a, b;
            memcpy(&a, p, 4), memcpy(&b, p + 4, 4);
            if (h.swap) a = bswap32(a), b = bswap32(b);
            eg = (int32_t)a, f = (int32_t)b;
        }
        NEKB_REQUIRE(eg >= 1 && eg <= h.nelgt && f >= 1 && f <= 12, ".re2: bad curved-side record");
        const int64_t slot = (eg - 1) * 12 + (f - 1);
        if (curve)
            for (int k = 0; k < 5; k++) curve[slot * 5 + k] = re2_word(p + (size_t)(2 + k) * h.wdsize, h.wdsize, h.swap);
        if (ccurve) ccurve[slot] = (char)p[(size_t)7 * h.wdsize];
    }
}

// ---------------------------------------------------------------------------------------------------------- .ma2
struct Ma2Header {
    int64_t nel = 0, nactive = 0, depth = 0, d2 = 0, npts = 0, nrank = 0, noutflow = 0;
    bool swap = false;
};

// core/map2.f:755-830: 132-byte header '#v001' + 7 integers (only neli, nnzi are read by the reference), endian tag,
// then per element 1 + 2^ldim int32: RSB leaf (processor id on the finest binary tree) and the vertex ids in SYMMETRIC
// corner order.
inline Ma2Header ma2_header(FileHandle &fh)
{
    char hdr[133] = {0};
    fh.read_at(0, hdr, 132, ".ma2 header");
    NEKB_REQUIRE(!strncmp(hdr, "#v001", 5), ".ma2: unknown header version");
    Ma2Header h;
    long long v[7] = {0, 0, 0, 0, 0, 0, 0};
    const int got = sscanf(hdr + 5, "%lld %lld %lld %lld %lld %lld %lld", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6]);
    NEKB_REQUIRE(got >= 2, ".ma2: cannot parse the header");
    h.nel = v[0], h.nactive = v[1], h.depth = v[2], h.d2 = v[3], h.npts = v[4], h.nrank = v[5], h.noutflow = v[6];
    float test;
    fh.read_at(132, &test, 4, ".ma2 endian tag");
    h.swap = endian_tag_swapped(test, ".ma2");
    return h;
}

inline void ma2_read(FileHandle &fh, const Ma2Header &h, int nlv, int64_t e0, int64_t nel, int32_t *leaf, int64_t *vertex)
{
    NEKB_REQUIRE(e0 >= 0 && nel >= 0 && e0 + nel <= h.nel, ".ma2: element range outside the file");
    const int64_t rec = (int64_t)(1 + nlv) * 4;
    std::vector<uint32_t> buf((size_t)(nel * (1 + nlv)));
    if (nel) fh.read_at(136 + e0 * rec, buf.data(), buf.size() * 4, ".ma2 records");
    for (int64_t e = 0; e < nel; e++) {
        const uint32_t *p = buf.data() + e * (1 + nlv);
        auto w = [&](int k) { return (int32_t)(h.swap ? bswap32(p[k]) : p[k]); };
        if (leaf) leaf[e] = w(0);
        if (vertex)
            for (int k = 0; k < nlv; k++) vertex[e * nlv + k] = (int64_t)w(1 + k);  // icopy48, map2.f:908
    }
}

// ---------------------------------------------------------------------------------------------------------- .co2
// core/map2.f:338-473 read_con: the connectivity file a parRSB build reads instead of .ma2 -- 132-byte header
// ('#v001' + 3 x i12, or '#v002' list-directed: nelgt, nelgv, nv), endian tag, then per element 1 + nv int32: the global
// element id and its vertex ids (consumed at map2.f:206-232: eid8 = wk4(ii+1), vtx8 = icopy48(wk4(ii+2..))).
struct Co2Header {
    int64_t nelgt = 0, nelgv = 0;
    int nv = 0;
    bool swap = false;
};

inline Co2Header co2_header(FileHandle &fh)
{
    char hdr[133] = {0};
    fh.read_at(0, hdr, 132, ".co2 header");
    NEKB_REQUIRE(!strncmp(hdr, "#v001", 5) || !strncmp(hdr, "#v002", 5), ".co2: unknown header version");
    Co2Header h;
    long long a = 0, b = 0, c = 0;
    const int got = sscanf(hdr + 5, "%lld %lld %lld", &a, &b, &c);
    NEKB_REQUIRE(got == 3, ".co2: cannot parse the header");
    h.nelgt = a, h.nelgv = b, h.nv = (int)c;
    NEKB_REQUIRE(h.nelgt >= 0 && h.nelgv >= 0 && h.nelgv <= h.nelgt && (h.nv == 4 || h.nv == 8), ".co2: bad header values");
    float test;
    fh.read_at(132, &test, 4, ".co2 endian tag");
    h.swap = endian_tag_swapped(test, ".co2");
    return h;
}

inline void co2_read(FileHandle &fh, const Co2Header &h, int64_t e0, int64_t nel, int64_t *eid, int64_t *vertex)
{
    NEKB_REQUIRE(e0 >= 0 && nel >= 0 && e0 + nel <= h.nelgt, ".co2: element range outside the file");
    const int nv = h.nv;
    const int64_t rec = (int64_t)(1 + nv) * 4;
    std::vector<uint32_t> buf((size_t)(nel * (1 + nv)));
    if (nel) fh.read_at(136 + e0 * rec, buf.data(), buf.size() * 4, ".co2 records");
    for (int64_t e = 0; e < nel; e++) {
        const uint32_t *p = buf.data() + e * (1 + nv);
        auto w = [&](int k) { return (int32_t)(h.swap ? bswap32(p[k]) : p[k]); };
        if (eid) eid[e] = (int64_t)w(0);
        if (vertex)
            for (int k = 0; k < nv; k++) vertex[e * nv + k] = (int64_t)w(1 + k);
    }
}

// core/math.f isort(a,ind,n): the reference's index heap sort (ascending, NOT stable -- the order of equal keys decides
// which elements land on either side of a partition boundary, so it is restated exactly).  1-based logic on 0-based storage.
inline void nek_isort(std::vector<int> &a, std::vector<int> &ind)
{
    const int n = (int)a.size();
    ind.resize(n);
    for (int j = 0; j < n; j++) ind[j] = j + 1;
    if (n <= 1) return;
    int L = n / 2 + 1, ir = n;
    for (;;) {
        int aa, ii;
        if (L > 1) {
            L = L - 1;
            aa = a[L - 1];
            ii = ind[L - 1];
        } else {
            aa = a[ir - 1];
            ii = ind[ir - 1];
            a[ir - 1] = a[0];
            ind[ir - 1] = ind[0];
            ir = ir - 1;
            if (ir == 1) {
                a[0] = aa;
                ind[0] = ii;
                return;
            }
        }
        int i = L, j = L + L;
        while (j <= ir) {
            if (j < ir && a[j - 1] < a[j]) j = j + 1;
            if (aa < a[j - 1]) {
                a[i - 1] = a[j - 1];
                ind[i - 1] = ind[j - 1];
                i = j;
                j = j + j;
            } else
                j = ir + 1;
        }
        a[i - 1] = aa;
        ind[i - 1] = ii;
    }
}

// core/map2.f:943-1026 assign_gllnid: element -> rank from the RSB leaves.  leaf[] is overwritten with gllnid (0-based ranks).
inline void assign_gllnid(std::vector<int> &gllnid, int64_t nelgt, int64_t nelgv, int np)
{
    NEKB_REQUIRE(np >= 1 && (int64_t)gllnid.size() >= nelgt, "assign_gllnid: bad arguments");
    int log2p = 0;
    while ((1 << (log2p + 1)) <= np) log2p++;
    const int np2 = 1 << log2p;
    auto vmax = [&](int64_t a, int64_t b) {
        int m = gllnid[a];
        for (int64_t e = a; e < b; e++) m = gllnid[e] > m ? gllnid[e] : m;
        return m;
    };
    if (np2 == np && nelgv == nelgt) {  // :950-960
        const int npstar = vmax(0, nelgt) + 1, nnpstr = npstar / np;
        NEKB_REQUIRE(nnpstr >= 1, "assign_gllnid: more ranks than leaves in the map file");
        for (int64_t e = 0; e < nelgt; e++) gllnid[e] = gllnid[e] / nnpstr;
        return;
    }
    if (np2 == np) {  // :962-978 conjugate heat transfer
        int npstar = std::max(np, vmax(0, nelgv) + 1), nnpstr = npstar / np;
        for (int64_t e = 0; e < nelgv; e++) gllnid[e] = gllnid[e] / nnpstr;
        npstar = std::max(np, vmax(nelgv, nelgt) + 1), nnpstr = npstar / np;
        for (int64_t e = nelgv; e < nelgt; e++) gllnid[e] = gllnid[e] / nnpstr;
        return;
    }
    NEKB_REQUIRE(nelgv == nelgt, "Conjugate heat transfer requires P=power of 2.");
    const int nel = (int)(nelgt / np), nmod = (int)(nelgt % np), npp = np - nmod;  // :996-1022
    std::vector<int> a(gllnid.begin(), gllnid.begin() + nelgt), ind;
    nek_isort(a, ind);
    int64_t k = 0;
    for (int ip = 0; ip < npp; ip++)
        for (int e = 0; e < nel; e++) a[k++] = ip;
    if (nmod > 0)
        for (int ip = npp; ip < np; ip++)
            for (int e = 0; e <= nel; e++) a[k++] = ip;
    // iswapt_ip (core/math.f): undo the permutation, x(ind(k)) = x_sorted(k)
    for (int64_t q = 0; q < nelgt; q++) gllnid[ind[q] - 1] = a[q];
}

}  // namespace nekb
