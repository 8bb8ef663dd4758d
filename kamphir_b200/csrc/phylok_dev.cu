// Device side of libphylok.so: the node-pair DP of PhyloKernel.kernel (phyloK2.py:158-209) as a height-level
// wavefront on sm_100a, plus the batched drivers (pair list, sim-vs-obs batch, Gram tiles).
//
//   C(i,j) = lambda * exp(-((a0-b0)^2 + (a1-b1)^2) / tau) * prod_{s=0,1} f_s          (phyloK2.py:184-199)
//   f_s    = 1                       if the slot-s children have different productions
//          = sigma + lambda          if both are tips
//          = sigma + C(ci_s, cj_s)   otherwise
//   K      = sum of C over production-matched internal-node pairs                      (phyloK2.py:203-204)
//
// One CTA evaluates one tree pair (persistent CTAs pull pairs from an atomic work counter).  The row tree's ROW part and
// the column tree's COLUMN part are staged into shared memory with one TMA bulk copy each (cp.async.bulk + mbarrier).
// Rows are sorted by (class, level, post-order), columns by (class, parent code, ancestors' parent codes, post-order):
// the matched cells of a pair form one dense cntA[c] x cntB[c] rectangle per production class c, level L of the
// wavefront is a contiguous row range of each rectangle, and the cells of a row that are ever read again (same parent
// code) are one contiguous column group -- only that group is stored (R ~ 0.39 M cells).  Thread t owns column t of
// every rectangle (column-stationary; labelled trees deal the 32-column chunks of a level's many narrow rectangles to
// the warps instead), keeps that node's data in registers while it walks the rows of a (level, class) rectangle two (or
// four) at a time, and all warps meet at one __syncthreads per level (child cells live in strictly lower levels).
// The branch lengths are pre-scaled per pair so that the squared distance is the exponent of the Gaussian in binary
// octaves (fp64: table + degree-4 polynomial + integer exponent add, 8 fp64 instructions; fp32: one MUFU.EX2).
// Stored cells live in shared memory for tiny pairs, otherwise in a per-CTA slab in HBM that stays L1/L2-resident:
// coalesced group writes, software-pipelined child-cell gathers (fp32 and labelled trees: the pipeline also runs past the
// end of a rectangle, PK_OVERRUN).  A pair is many short rectangles (~27 with gathers per 500-tip pair), so the code between
// the row loops counts: per-level descriptors carry the addresses a thread needs to set up its column (PK_DESC2).
// The kernel is instantiated per CTA-size tier (64 ... 1024 threads), each with the
// register budget, resident CTAs per SM and rows in flight that measured best (table in DESIGN.md section 4).
// No tensor cores (not a contraction), no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include "phylok_internal.h"

#define PK_CUDA(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return pk_fail(PK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// work description
// ---------------------------------------------------------------------------------------------------
struct DpArgs {
    const unsigned char *blocksA;   // packed tree blocks of forest A
    const int4 *metaA;              // per tree two records: row part, column part: {offset / 16, bytes, n, nlev}
    const unsigned char *blocksB;
    const int4 *metaB;
    int mode;                       // 0: explicit pair list, 1: Gram tiles, 2: diagonal (i,i)
    const int2 *pairs;              // mode 0
    const int2 *tiles;              // mode 1: tile coordinates (ti, tj), tj >= ti
    int tile_shift;                 // log2(tile edge)
    int n_trees;                    // mode 1/2
    long long total;                // work items
    double *out;                    // mode 0: out[w]; mode 1: out[i*ld+j] and out[j*ld+i]; mode 2: out[i]
    long long ld;
    const double *diag;             // mode 1, optional: write K / sqrt(diag[i]*diag[j])
    double lambda, neg_inv_tau, sigma;
    void *scratch;                  // HBM path: per-CTA C slabs
    long long slab_elems;
    unsigned long long *counter;    // work counter
    unsigned long long *nonfinite;  // per-context count of results that came out inf / NaN (fp32 cells overflow beyond 3e38)
    int tree_region;                // shared-memory bytes reserved for the staged ROW part
    int col_region;                 // ... and for the staged COLUMN part
    int rowinfo_bytes;              // row offsets
    unsigned char *stage;           // STAGED = false: per-CTA staging area in global memory (tree parts + row offsets)
    long long stage_stride;
    double bscale;                  // sqrt(64 / (ln2 tau)) (fp64) or sqrt(1 / (ln2 tau)) (fp32), applied to the staged branch lengths
    int clamp_hi;                   // fp64 Gaussian: clamp of the exponent's high word
    unsigned long long exp_tab[64]; // fp64 Gaussian: bits of lambda * 2^(j/64) with the high word biased by -(j << 14)
};

// ---------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// ---------------------------------------------------------------------------------------------------
// lambda * exp(-d^2 / tau), U independent values at a time
// ---------------------------------------------------------------------------------------------------
// fp64: the branch lengths are pre-scaled by sqrt(64 / (ln2 tau)) when a pair is staged, so the squared distance E of two
// scaled length pairs is the exponent in units of 1/64 octave: lambda * exp(-d^2/tau) = lambda * 2^(-E/64).
//   t = MAGIC - E (MAGIC = 1.5 * 2^52)  ->  low word of t = k = -round(E);  r = E - round(E), |r| <= 1/2
//   lambda * 2^(-E/64) = [lambda * 2^(j/64)] * 2^e * 2^(-r/64),   j = k & 63, e = k >> 6
// The bracket comes from a 64-entry table (filled by the host in fp64), 2^(-r/64) from a degree-4 near-minimax
// polynomial (max relative error 5e-15, checked on the host over 4e6 arguments), and 2^e is one integer add on the
// table value's high word: the table stores hi - (j << 14), so hi + (k << 14) adds e << 20.  8 fp64-pipe instructions
// and 5 integer ones.  E is clamped (one unsigned min on its high word) so that the result stays a normal number
// (~1e-307 instead of 0 beyond exp(-700)).  The table is replicated 16x in shared memory (entry j of lane l at word
// 16 j + (l & 15)): a 64-bit LDS is served per half-warp, so every lookup is conflict-free.  The U chains are written
// stage by stage so the scheduler interleaves them.
#define PK_EXP_TAB 64
#ifndef PK_WHATIF
#define PK_WHATIF 0      // timing experiments (wrong results), never set in a shipped build
#endif
#ifndef PK_TAB_MAD
#define PK_TAB_MAD 1
#endif
#define PK_EXP_REP 16
__constant__ double pk_poly[4] = {-0x1.62e42fefa2417p-7, 0x1.ebfbdff82bb76p-15, -0x1.c6b0b9214b9aep-23,
                                  0x1.3b2acb2c36b23p-31};   // 2^(-r/64) = 1 + r (c1 + c2 r + c3 r^2 + c4 r^3)
template <typename T>
struct GaussK {
    uint32_t tab_addr;   // shared-memory address of this lane's copy of table entry 0
    int clamp_hi;        // high word of the largest E that keeps the result normal
    T lambda, neg_inv_tau;
    uint32_t extra_addr; // shared-memory address of {sigma, -, c3, c4} behind the table
    bool regc;           // pk_rect2: the two highest polynomial coefficients come in registers (c3, c4), read from shared memory
    double c3, c4;       // with a volatile load so that ptxas cannot re-read them from the constant bank in every iteration
};
template <int U>
__device__ __forceinline__ void pk_gauss(const double (&E)[U], double (&out)[U], const GaussK<double> &gk) {
    double Ec[U], t[U], r[U], q[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        Ec[u] = __hiloint2double((int)min((unsigned)__double2hiint(E[u]), (unsigned)gk.clamp_hi), __double2loint(E[u]));
#pragma unroll
    for (int u = 0; u < U; ++u) t[u] = 6755399441055744.0 - Ec[u];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = (t[u] - 6755399441055744.0) + Ec[u];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = gk.regc ? fma(gk.c4, r[u], gk.c3) : fma(pk_poly[3], r[u], pk_poly[2]);
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = fma(q[u], r[u], pk_poly[1]);
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = fma(q[u], r[u], pk_poly[0]);
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = fma(q[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int k = __double2loint(t[u]);
        unsigned lo;
        int hi;
#if PK_TAB_MAD           // entry address as AND + multiply-add (LOP3, IMAD) instead of the compiler's SHL, LOP3, IADD
        uint32_t ta;
        asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, 63;\n\tmad.lo.u32 %0, t, 128, %2;\n\t}" : "=r"(ta) : "r"(k), "r"(gk.tab_addr));
        if (PK_WHATIF & 4) { asm("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(ta)); lo = 0; } else
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(ta));
#else
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(gk.tab_addr + ((k & (PK_EXP_TAB - 1)) << 7)));
#endif
        out[u] = q[u] * __hiloint2double(hi + (k << 14), (int)lo);
    }
}
// fp32: the lengths are pre-scaled by sqrt(log2(e) / tau), so lambda * exp(-d^2/tau) = lambda * 2^(-E): one MUFU.EX2
template <int U>
__device__ __forceinline__ void pk_gauss(const float (&E)[U], float (&out)[U], const GaussK<float> &gk) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-E[u]));
        out[u] = gk.lambda * e;
    }
}

// integer division of small non-negative ints through fp32 (exact while a < 2^22 and the quotient's fractional part
// stays away from 0/1 by 0.5/b, which the +0.5 guarantees); the kernel's uniform per-level bookkeeping uses it
__device__ __forceinline__ int pk_div(int a, int b) { return __float2int_rz(__fdividef((float)a + 0.5f, (float)b)); }

// per row of the row tree: branch lengths, per child slot the offset of the child's stored row in C, the offset that
// turns a column position into this row's stored cell, and three packed bytes: class + 1 of child 0, of child 1
// (0 = tip) and the row's parent code
template <typename T>
struct __align__(16) RowRec {
    T a0, a1;
    int st2, pk2;                       // fp32, PK_PACKQ: copies of st_off / packed, so one LDS.128 brings the lengths and what the store needs
    int off0, off1, st_off, packed;     // bytes 16..31: read as one int4 (PK_PACKQ: the first 8 bytes, with the child classes in the top bytes)
};
template <>
struct __align__(16) RowRec<double> {
    double a0, a1;
    int off0, off1, st_off, packed;
};

// one (level, class) rectangle of the wavefront, precomputed per pair so the warps only read it.  In the PK_DESC2 kernels
// (see the wavefront) r0 and colbase hold the shared-memory addresses of the rectangle's first column in the b0 and code0
// arrays; for unlabelled trees chunk_end / nchunks hold the distances to the b1 and to the code1 / parent-code arrays and
// bits 12-13 of the flags the slots to the next live descriptor.
struct __align__(16) LevelDesc {
    int r0;         // first row (kernel order) of the rectangle
    int rows;       // rows at this level in this class (0: empty, skipped)
    int wB;         // columns = nodes of the column tree in this class
    int colbase;    // first column (kernel order in the column tree)
    int chunk_end;  // DEAL: the level's 32-column chunks are numbered across its rectangles, this one owns [chunk_end - n, chunk_end)
    int flags;      // bit 8: last rectangle of its level (a barrier follows); bit 9 / 10: slot 0 / 1 always a tip
    int nchunks;    // DEAL, in the level's first descriptor: chunks of the whole level
    int pad1;
};
#ifndef PK_DESC_MAX
#define PK_DESC_MAX 64
#endif
static_assert(PK_DESC_MAX >= PK_MAX_CLASSES, "one level's descriptors (one per class) must fit the descriptor buffer");

// ---------------------------------------------------------------------------------------------------
// stored C cells of one pair: a shared-memory array (tiny trees) or this CTA's slab in HBM.  The slab base is an opaque
// 64-bit value in a vector register pair, so every cell address is one IMAD.WIDE with an immediate 8 (or 4); the accesses
// are explicit ld.global / st.global (the opaque base would otherwise decay to generic LD/ST).
// ---------------------------------------------------------------------------------------------------
// -DPK_DEBUG_BOUNDS (scripts/bounds_check.sh; compute-sanitizer is not available on the GPU pool): every cell index is
// checked against the slab / shared-memory array size, violations are counted and the launch fails.
#ifndef PK_ST_NOCLOBBER
#define PK_ST_CLOBBER : "memory"
#else
#define PK_ST_CLOBBER
#endif
template <typename T, bool CSMEM>
struct Cells {
    T *sm;
    unsigned long long base;
#ifdef PK_DEBUG_BOUNDS
    unsigned limit;
    unsigned long long *err;
    __device__ __forceinline__ int chk(int idx) const {
        if ((unsigned)idx >= limit) {
            atomicAdd(err, 1ull);
            return 0;
        }
        return idx;
    }
#else
    __device__ __forceinline__ int chk(int idx) const { return idx; }
#endif
    __device__ __forceinline__ T ld(int idx) const {
        idx = chk(idx);
        if (CSMEM) return sm[idx];
        T v;
        const unsigned long long addr = base + (unsigned long long)(unsigned)idx * sizeof(T);
        if (sizeof(T) == 8) asm volatile("ld.global.f64 %0, [%1];" : "=d"(*reinterpret_cast<double *>(&v)) : "l"(addr));
        else asm volatile("ld.global.f32 %0, [%1];" : "=f"(*reinterpret_cast<float *>(&v)) : "l"(addr));
        return v;
    }
    __device__ __forceinline__ void st(int idx, T v) const {
        idx = chk(idx);
        if (CSMEM) {
            sm[idx] = v;
            return;
        }
        const unsigned long long addr = base + (unsigned long long)(unsigned)idx * sizeof(T);
        if (sizeof(T) == 8) asm volatile("st.global.f64 [%0], %1;" ::"l"(addr), "d"(*reinterpret_cast<double *>(&v)) PK_ST_CLOBBER);
        else asm volatile("st.global.f32 [%0], %1;" ::"l"(addr), "f"(*reinterpret_cast<float *>(&v)) PK_ST_CLOBBER);
    }
};

// ---------------------------------------------------------------------------------------------------
// one column of one (level, class) rectangle: the thread keeps its column's node in registers and walks the rows U at a
// time.  G0 / G1: slot 0 / 1 of the rows has internal children somewhere in this class (otherwise that slot is a tip in
// every row and its factor is a per-column constant: sigma + lambda against a tip, 1 against an internal node).
//   C(i,j) = lambda * exp(-((a0-b0)^2 + (a1-b1)^2) / tau) * f0 * f1                       (phyloK2.py:184-199)
// Software pipeline: the second half of the NEXT group's row records (child-row offsets, class bytes) is read and its
// child-cell gathers are issued before the Gaussians of the current group, so the L2 / HBM latency of the gathers is
// covered.  (Measured and dropped: carrying the store offsets in registers instead of re-reading them, predicated
// gathers that skip the mismatching lanes, gathers two groups ahead: all slower, the loop is sensitive to instruction
// count and registers.)  Returns the column's partial sum of C over the rows.
// ---------------------------------------------------------------------------------------------------
template <bool STAGED>
__device__ __forceinline__ int4 ld_int4(const int4 *p) {
    int4 v;
    if (STAGED)      // always one LDS.128 (never split into scalar loads)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    else             // parts staged in global memory (trees beyond the shared-memory limit): one uniform 16-byte load, L1-resident
        asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// PK_OVERRUN (HBM slabs only): the software pipeline always reads the NEXT U row records and issues their gathers, also
// when they lie past the rectangle -- no pointer clamp in the loop, and the last 1..U-1 rows of a rectangle find their
// child cells already in flight.  What makes it safe: the records after a rectangle are other rows of the same staged
// tree (patched, so their offsets are real stored-row offsets), U never-matching records follow the last row (written
// after staging), and the slab has a slack of the largest column group, so offset + rank stays inside the slab; the
// values fetched for rows beyond the rectangle are never used.
#ifndef PK_OVERRUN
#define PK_OVERRUN 1
#endif
#ifndef PK_OPAQUE_Q
#define PK_OPAQUE_Q 1
#endif
#ifndef PK_FENCE_ALL
#define PK_FENCE_ALL 0
#endif
#define PK_PAD_RECORDS 4      // >= the largest U
// PK_PACKQ: the class of a row's child (+ 1, 0 = tip) rides in the top byte of the child's stored-row offset, so the
// prefetch of a row needs 8 bytes of its record instead of 16 (one shared-memory wavefront instead of two per row and
// warp; the LSU data pipe is the busiest unit of the kernel).  The column keeps the same byte in the top byte of its test
// word and its rank pre-biased by minus that byte, so "offset word + biased rank" is the cell index without any masking.
// Needs offsets below 2^24 (pairs up to ~10 000 tips); bigger slabs take the plain records.
#ifndef PK_PACKQ
#define PK_PACKQ 1
#endif
#ifndef PK_QMASK
#define PK_QMASK 0x7f000000
#endif
template <bool STAGED>
__device__ __forceinline__ int2 ld_int2(const int2 *p) {
    int2 v;
    if (STAGED) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)));
    else asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// child cell of slot 0 / 1: the stored cell when the children's classes agree, else cell 1 (= 1, productions differ)
// (PQ: the mask leaves bit 31 out -- class + 1 <= 65 needs 7 bits -- so that the test cannot be rewritten as an unsigned
// compare of the xor, which costs two instructions; as it stands it is one LOP3 with a predicate result)
#define PK_GATHER0(q) (PQ ? C.ld((((q).x ^ qb0) & PK_QMASK) == 0 ? (q).x + kb0 : 1) : C.ld((((q).w ^ qb0) & 0xff) == 0 ? (q).x + kb0 : 1))
#define PK_GATHER1(q) (PQ ? C.ld((((q).y ^ qb1) & PK_QMASK) == 0 ? (q).y + kb1 : 1) : C.ld((((q).w ^ qb1) & 0xff00) == 0 ? (q).y + kb1 : 1))
// the part of a row record the gathers need: PQ: {offset | class << 24} of both children (8 bytes), else the whole second half
#define PK_LOADQ(ptr)                                                                                   \
    int4 q;                                                                                             \
    if (PQ) {                                                                                           \
        const int2 q2 = ld_int2<STAGED>(reinterpret_cast<const int2 *>(ptr));                           \
        q = make_int4(q2.x, q2.y, 0, 0);                                                                \
    } else {                                                                                            \
        q = ld_int4<STAGED>(ptr);                                                                       \
    }
template <typename T, bool CSMEM, bool STAGED, bool G0, bool G1, int U, bool FENCE, bool OVRQ, bool PQ>
__device__ __forceinline__ T pk_rect(const RowRec<T> *rp, int rows, int jb, T b0, T b1, int cb0, int cb1, int pkey,
                                     const Cells<T, CSMEM> &C, const GaussK<T> &gk, T sigma, T tipf) {
    constexpr bool OVR = OVRQ && !CSMEM;         // the prefetch may run past the rectangle (see PK_OVERRUN)
    // child class + 1 (0 = tip) where the row record has it: top byte of the offset word (PQ) or byte 0 / 1 of rec.packed
    int qb0 = PQ ? (int)((unsigned)cb0 & 0xff000000u) : (int)((unsigned)cb0 >> 24);
#if PK_OPAQUE_Q
    asm("" : "+r"(qb0));      // hide the value range: the class test of slot 0 stays one LOP3 with a predicate result
#endif
    const int qb1 = PQ ? (int)((unsigned)cb1 & 0xff000000u) : (int)(((unsigned)cb1 >> 24) << 8);
    // rank in its group; PQ: minus the class byte, so that offset word + rank is the cell index
    const int kb0 = PQ ? (cb0 & 0xffffff) - (int)((unsigned)cb0 & 0xff000000u) : (cb0 & 0xffffff);
    const int kb1 = PQ ? (cb1 & 0xffffff) - (int)((unsigned)cb1 & 0xff000000u) : (cb1 & 0xffffff);
    T cf = (T)1;                                                           // factor of the slots that are tips in every row
    if (!G0) cf = qb0 == 0 ? tipf : (T)1;
    if (!G1) cf *= qb1 == 0 ? tipf : (T)1;
    T part = (T)0;
    T f0[U], f1[U];
    int rr = 0;
    const int4 *ip = reinterpret_cast<const int4 *>(rp);                   // ip[2 r + 1] = {off0, off1, st_off, packed}
    if ((G0 || G1) && U <= rows) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            PK_LOADQ(ip + 2 * u + 1)
            if (G0) f0[u] = PK_GATHER0(q);
            if (G1) f1[u] = PK_GATHER1(q);
        }
    }
    for (; rr + U <= rows; rr += U, rp += U, ip += 2 * U) {
        T x[U], ev[U], g[U];
        int so[U], pk[U];      // fp32 + PQ: the store offset and packed word arrive with the lengths
#pragma unroll
        for (int u = 0; u < U; ++u) {
            T a0, a1;
            if (PQ && sizeof(T) == 4) {
                const int4 h = ld_int4<STAGED>(ip + 2 * u);
                a0 = (T)__int_as_float(h.x);
                a1 = (T)__int_as_float(h.y);
                so[u] = h.z;
                pk[u] = h.w;
            } else {
                a0 = rp[u].a0;
                a1 = rp[u].a1;
            }
            const T d0 = a0 - b0, d1 = a1 - b1;
            x[u] = fma(d1, d1, d0 * d0);
            g[u] = G0 ? (G1 ? f0[u] * f1[u] : f0[u] * cf) : (G1 ? f1[u] * cf : cf);
        }
        if (G0 || G1) {   // always prefetch (the last group re-reads itself: no join, no register copies)
            const int4 *np = OVR || rr + 2 * U <= rows ? ip + 2 * U : ip;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                PK_LOADQ(np + 2 * u + 1)
                if (G0) f0[u] = PK_GATHER0(q);
                if (G1) f1[u] = PK_GATHER1(q);
            }
            // ptxas is free to sink the gathers below the Gaussians, and whether it does depends on unrelated code around
            // the loop and on the register budget (375 instead of 445 Gcells/s when the chunk-dealing level loop was
            // introduced).  Memory operations do not cross a warp barrier: FENCE pins the gathers here (costs 2-4 % where
            // ptxas behaves by itself).  Measured and dropped: touching the gathered factors only after the Gaussian
            // (v = (e * f0) * f1) with the next gathers issued after that -- ptxas pairs each row's use with its own
            // Gaussian, the distance from gather to first use stays ~30 instructions (449 against 460).
            // The mask is taken here, not once per rectangle: lanes that are not converged at this point (a build with
            // divergent code in the loop, e.g. PK_DEBUG_BOUNDS) each synchronise with whoever is with them; a mask taken at
            // the rectangle's entry deadlocked exactly there (labelled trees, one-warp CTA).
            if (FENCE) __syncwarp(__activemask());
        }
        pk_gauss<U>(x, ev, gk);
#pragma unroll
        for (int u = 0; u < U; ++u) {     // e * g feeds two FMAs: the stored cell sigma + C and the running sum
            if (!(PQ && sizeof(T) == 4)) {
                pk[u] = rp[u].packed;
                so[u] = rp[u].st_off;
            }
            if (((pk[u] ^ pkey) & 0xff0000) == 0) C.st(so[u] + jb, fma(ev[u], g[u], sigma));
            part = fma(ev[u], g[u], part);
        }
    }
    // the last 1..U-1 rows: all their gathers first (one exposed round trip, not one per row); with OVR the last group of
    // the loop above has already fetched them
    const int rem = rows - rr;
    if (U > 1 && rem > 0) {
        if ((G0 || G1) && !(OVR && rr > 0)) {
#pragma unroll
            for (int u = 0; u < U - 1; ++u) {
                if (u < rem) {
                    PK_LOADQ(ip + 2 * u + 1)
                    if (G0) f0[u] = PK_GATHER0(q);
                    if (G1) f1[u] = PK_GATHER1(q);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U - 1; ++u) {
            if (u < rem) {
                const T d0 = rp[u].a0 - b0, d1 = rp[u].a1 - b1;
                T x1[1] = {fma(d1, d1, d0 * d0)}, e1[1];
                pk_gauss<1>(x1, e1, gk);
                const T g = G0 ? (G1 ? f0[u] * f1[u] : f0[u] * cf) : (G1 ? f1[u] * cf : cf);
                if (((rp[u].packed ^ pkey) & 0xff0000) == 0) C.st(rp[u].st_off + jb, fma(e1[0], g, sigma));
                part = fma(e1[0], g, part);
            }
        }
    }
    return part;
}

// ---------------------------------------------------------------------------------------------------
// pk_rect2: the same rectangle walk for the unlabelled fp64 tiers with both tree parts in shared memory and C in an HBM
// slab, written against instruction count (round 2, second half).
//  * The row records are addressed by their 32-bit shared-memory address and the loop runs on that address alone (end
//    test = one compare, prefetch clamp = one minimum).
//  * Every cell address is ONE IMAD.WIDE.U32 onto a per-lane 64-bit base that already holds the lane's part of the index.
//    The cell indices of such a kernel start at PK_XOFF instead of 0 (cells.base is moved down by as much), the patched
//    record words carry a tag in their top byte -- class + 1 of the child in the two offset words, the parent code in the
//    store word -- and the lane's bases are moved down by its own tag:
//      gather slot s :  bk_s = slab + 8 * (rank_s - tag_s),  word + rank_s - tag_s = the child's cell when the tags agree;
//                       otherwise the lane selects alt_s = tag_s + (PK_XOFF + 1 - rank_s), which leads from bk_s to the cell
//                       that holds 1 -- and whose top byte is tag_s again (PK_XOFF + 1 >= any rank), so alt_s is also the
//                       word the class test compares with: LOP3 (predicate), SEL, IMAD.WIDE, LDG per gather
//      store         :  bj = slab + 8 * (column - tag), word + column - tag = the stored cell: LOP3, IMAD.WIDE, STG
// ---------------------------------------------------------------------------------------------------
#ifndef PK_RECT2
#define PK_RECT2 2      // 0: off, 1: unlabelled fp64 kernels only, 2: every kernel with staged tree parts and C in HBM slabs
#endif
#ifndef PK_R2_FENCE
#define PK_R2_FENCE 0
#endif
#ifndef PK_R2_ROT
#define PK_R2_ROT 1
#endif
#ifndef PK_DESC2L
#define PK_DESC2L 1
#endif
#ifndef PK_DESC2
#define PK_DESC2 1      // unlabelled pk_rect2 kernels: a rectangle's descriptor carries the addresses of its columns' data
#endif
#define PK_XOFF 16384          // >= the largest rank of a node in its (class, parent code) group; offsets stay below 2^24
__device__ __forceinline__ void lds_f64x2(uint32_t a, double &x, double &y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ int lds_i32(uint32_t a) {
    int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_i32x2(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds_i32x4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
template <typename T>
__device__ __forceinline__ T lds_real(uint32_t a) {
    T v;
    if (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(*reinterpret_cast<double *>(&v)) : "r"(a));
    else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(*reinterpret_cast<float *>(&v)) : "r"(a));
    return v;
}
template <typename T>
__device__ __forceinline__ unsigned long long cell_addr(unsigned long long b, int idx) {     // one IMAD.WIDE.U32
    return b + (unsigned long long)(unsigned)idx * (unsigned long long)sizeof(T);
}
#ifdef PK_DEBUG_BOUNDS
// pk_rect2 addresses cells through per-lane bases, so the check is on the final address: it must lie inside this CTA's slab
// (the kernel records the slab's bounds and the violation counter per CTA before the first pair)
__device__ unsigned long long pk_dbg_lo[4096], pk_dbg_hi[4096];
__device__ unsigned long long *pk_dbg_err[4096];
template <typename T>
__device__ __forceinline__ bool pk_dbg_ok(unsigned long long addr) {
    const unsigned bx = blockIdx.x & 4095u;
    if (addr >= pk_dbg_lo[bx] && addr + sizeof(T) <= pk_dbg_hi[bx] && addr % sizeof(T) == 0) return true;
    atomicAdd(pk_dbg_err[bx], 1ull);
    return false;
}
#endif
template <typename T>
__device__ __forceinline__ T ldg_cell(unsigned long long b, int idx) {
    T v;
#ifdef PK_DEBUG_BOUNDS
    if (!pk_dbg_ok<T>(cell_addr<T>(b, idx))) return (T)1;
#endif
    if (sizeof(T) == 8) asm volatile("ld.global.f64 %0, [%1];" : "=d"(*reinterpret_cast<double *>(&v)) : "l"(cell_addr<T>(b, idx)));
    else asm volatile("ld.global.f32 %0, [%1];" : "=f"(*reinterpret_cast<float *>(&v)) : "l"(cell_addr<T>(b, idx)));
    return v;
}
template <typename T>
__device__ __forceinline__ void stg_cell(unsigned long long b, int idx, T v) {
#ifdef PK_DEBUG_BOUNDS
    if (!pk_dbg_ok<T>(cell_addr<T>(b, idx))) return;
#endif
    if (sizeof(T) == 8) asm volatile("st.global.f64 [%0], %1;" ::"l"(cell_addr<T>(b, idx)), "d"(*reinterpret_cast<double *>(&v)) : "memory");
    else asm volatile("st.global.f32 [%0], %1;" ::"l"(cell_addr<T>(b, idx)), "f"(*reinterpret_cast<float *>(&v)) : "memory");
}
#ifndef PK_R2_SIGMA_REG
#define PK_R2_SIGMA_REG 1
#endif
// WIDEP: parent codes beyond 127 are possible (labelled trees with 63 or 64 classes): the store test compares the whole top byte
// CONV: the column's words were converted when the pair was staged (PK_DESC2 kernels): cb0 / cb1 arrive as rank - tag (0 for a
// tip child, negative otherwise: a tag is at least 2^24, a rank below it) and pcode as the parent code << 24
template <typename T, bool G0, bool G1, int U, bool WIDEP, bool CONV>
__device__ __forceinline__ double pk_rect2(uint32_t ra, int rows, int jb, T b0, T b1, int cb0, int cb1, int pcode,
                                           unsigned long long base, const GaussK<T> &gk_in, T sigma, T tipf, double acc) {
    constexpr long long W = (long long)sizeof(T);
    const int d0 = CONV ? cb0 : (cb0 & 0xffffff) - (int)((unsigned)cb0 & 0xff000000u);
    const int d1 = CONV ? cb1 : (cb1 & 0xffffff) - (int)((unsigned)cb1 & 0xff000000u);
    const int tag0 = d0, tag1 = d1;      // only ever compared with 0 (tip child)
    unsigned long long bk0 = base + (unsigned long long)((long long)d0 * W);
    unsigned long long bk1 = base + (unsigned long long)((long long)d1 * W);
    int alt0 = (PK_XOFF + 1) - d0, alt1 = (PK_XOFF + 1) - d1;      // = tag + (PK_XOFF + 1 - rank)
    const unsigned ptag = CONV ? (unsigned)pcode : (unsigned)pcode << 24;
    constexpr unsigned PMASK = WIDEP ? 0xff000000u : (unsigned)PK_QMASK;
    unsigned long long bj = base + (unsigned long long)(((long long)jb - (long long)ptag) * W);
    GaussK<T> gk = gk_in;
#if PK_R2_SIGMA_REG
    // fp64: sigma and the two highest polynomial coefficients stay in register pairs: read from shared memory with volatile
    // loads, which ptxas cannot repeat (it otherwise re-reads them from the constant bank in every iteration)
    if constexpr (sizeof(T) == 8) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(*reinterpret_cast<double *>(&sigma)) : "r"(gk.extra_addr));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(gk.c3), "=d"(gk.c4) : "r"(gk.extra_addr + 16u));
        gk.regc = true;
    }
#endif
    // opaque and pinned here: kept in registers as they are, not re-derived per access
    if (G0) asm volatile("" : "+l"(bk0), "+r"(alt0));
    if (G1) asm volatile("" : "+l"(bk1), "+r"(alt1));
    asm volatile("" : "+l"(bj));
#if PK_WHATIF & 8      // timing experiment only (wrong results): no gathers at all
#define PK2_G0(w) (T)(w)
#define PK2_G1(w) (T)(w)
#elif PK_WHATIF & 1    // timing experiment only (wrong results): mismatching lanes skip the load
#define PK2_G0(w) ((((w) ^ alt0) & PK_QMASK) == 0 ? ldg_cell<T>(bk0, (w)) : f0[u])
#define PK2_G1(w) ((((w) ^ alt1) & PK_QMASK) == 0 ? ldg_cell<T>(bk1, (w)) : f1[u])
#else
#define PK2_G0(w) ldg_cell<T>(bk0, (((w) ^ alt0) & PK_QMASK) == 0 ? (w) : alt0)
#define PK2_G1(w) ldg_cell<T>(bk1, (((w) ^ alt1) & PK_QMASK) == 0 ? (w) : alt1)
#endif
    T cf = (T)1;                                                      // factor of the slots that are tips in every row
    if (!G0) cf = tag0 == 0 ? tipf : (T)1;
    if (!G1) cf *= tag1 == 0 ? tipf : (T)1;
    // fp64: the thread's running sum over the pair is passed through; fp32: an fp32 sum per rectangle, added to it at the end
    typename std::conditional<sizeof(T) == 8, double, float>::type part = sizeof(T) == 8 ? acc : 0.0;
    T f0[U], f1[U];
    const uint32_t rend = ra + 32u * (uint32_t)(rows & ~(U - 1));     // first record after the full groups
    const uint32_t rlast = rend - 32u * U;                            // the last full group
    // Rotated software pipeline: the Gaussians of a group are computed while its child cells are in flight -- also those of
    // the first group of the rectangle (the exposed round trip of every rectangle's head) --, and the loop needs no prefetch
    // clamp: it runs up to the last full group, which is finished behind it.
    // Row record: fp64 {a0, a1 | off0, off1, store word, -}; fp32 {a0, a1, store word, - | off0, off1, -, -} (one LDS.128
    // brings the lengths and the store word)
#define PK2_GATHER(at)                                           \
    if (G0 || G1) {                                              \
        _Pragma("unroll") for (int u = 0; u < U; ++u) {          \
            const int2 q = lds_i32x2((at) + 32u * u + 16);       \
            if (G0) f0[u] = PK2_G0(q.x);                         \
            if (G1) f1[u] = PK2_G1(q.y);                         \
        }                                                        \
    }
#define PK2_GAUSS(at, UU, ev_, sw_)                              \
    {                                                            \
        T x[UU];                                                 \
        _Pragma("unroll") for (int u = 0; u < UU; ++u) {         \
            T a0, a1;                                            \
            if constexpr (sizeof(T) == 8) {                      \
                lds_f64x2((at) + 32u * u, *reinterpret_cast<double *>(&a0), *reinterpret_cast<double *>(&a1));   \
            } else {                                             \
                const int4 h = lds_i32x4((at) + 32u * u);        \
                a0 = (T)__int_as_float(h.x);                     \
                a1 = (T)__int_as_float(h.y);                     \
                sw_[u] = h.z;                                    \
            }                                                    \
            const T d0 = a0 - b0, d1 = a1 - b1;                  \
            x[u] = fma(d1, d1, d0 * d0);                         \
        }                                                        \
        pk_gauss<UU>(x, ev_, gk);                                \
    }
#define PK2_FINISH(at, UU, ev_, g_, sw_)                         \
    _Pragma("unroll") for (int u = 0; u < UU; ++u) {             \
        const int sw = sizeof(T) == 8 ? lds_i32((at) + 32u * u + 24) : sw_[u];      \
        if ((((unsigned)sw ^ ptag) & PMASK) == 0) stg_cell<T>(bj, sw, fma(ev_[u], g_[u], sigma));   \
        part = fma(ev_[u], g_[u], part);                         \
    }
#define PK2_PROD(g_, UU)                                         \
    _Pragma("unroll") for (int u = 0; u < UU; ++u) g_[u] = G0 ? (G1 ? f0[u] * f1[u] : f0[u] * cf) : (G1 ? f1[u] * cf : cf);
    if (ra != rend) {
        T ev[U], g[U];
        int sw4[U];
#if PK_WHATIF & 32      // timing experiment only (wrong results): the first group of a rectangle takes no child cells
        _Pragma("unroll") for (int u = 0; u < U; ++u) { f0[u] = (T)(ra + u); f1[u] = (T)(rows + u); }
#else
        PK2_GATHER(ra)
#endif
        PK2_GAUSS(ra, U, ev, sw4)
        for (; ra != rlast; ra += 32u * U) {
            PK2_PROD(g, U)
            PK2_GATHER(ra + 32u * U)
            PK2_FINISH(ra, U, ev, g, sw4)
            PK2_GAUSS(ra + 32u * U, U, ev, sw4)
        }
        PK2_PROD(g, U)
        PK2_FINISH(ra, U, ev, g, sw4)
        ra += 32u * U;
    }
    const int rem = rows & (U - 1);
    if (U > 1 && rem > 0) {       // the last 1 .. U - 1 rows: all their gathers first, then one row at a time
        if (G0 || G1) {
#pragma unroll
            for (int u = 0; u < U - 1; ++u) {
                if (u < rem) {
                    const int2 q = lds_i32x2(ra + 32u * u + 16);
                    if (G0) f0[u] = PK2_G0(q.x);
                    if (G1) f1[u] = PK2_G1(q.y);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U - 1; ++u) {
            if (u < rem) {
                T e1[1], g1[1];
                int s1[1];
                const uint32_t rr = ra + 32u * u;      // (the macros have a loop variable u of their own)
                PK2_GAUSS(rr, 1, e1, s1)
                g1[0] = G0 ? (G1 ? f0[u] * f1[u] : f0[u] * cf) : (G1 ? f1[u] * cf : cf);
                PK2_FINISH(rr, 1, e1, g1, s1)
            }
        }
    }
#undef PK2_GATHER
#undef PK2_GAUSS
#undef PK2_FINISH
#undef PK2_PROD
#undef PK2_G0
#undef PK2_G1
    if constexpr (sizeof(T) == 8) return part;
    else return acc + (double)part;
}

// ---------------------------------------------------------------------------------------------------
// the DP kernel
// ---------------------------------------------------------------------------------------------------
#ifndef PK_UNROLL
#define PK_UNROLL 4
#endif
#ifndef PK_COLGLOBAL_DEFAULT
#define PK_COLGLOBAL_DEFAULT 0      // COLUMN part in the global staging area for the pairs that fit one CTA per SM otherwise
#endif
#ifndef PK_U2_BOUND
#define PK_U2_BOUND 352
#endif
#ifndef PK_USMALL
#define PK_USMALL 2
#endif
#ifndef PK_MINB
#define PK_MINB 3
#endif
// Column-stationary mapping: warp w owns columns [32w, 32w+32) of every class rectangle, a lane one column.  Every warp
// therefore walks the same rows of a (level, class) rectangle -- no work splitting, no per-warp bookkeeping, perfectly
// balanced level barriers -- and a CTA has as many warps as the widest class needs (6 for 500-tip trees), several CTAs
// (pairs) being resident per SM.
// DEAL: the 32-column chunks of all rectangles of a level are dealt to the warps (labelled trees: many narrow classes);
// otherwise thread t owns column t of every rectangle and the rectangles of a level run one after the other.
// STG: where a pair's tree parts are staged.
//   1  ROW and COLUMN part in shared memory (TMA bulk copies): every tier up to ~1 600 tips at several CTAs per SM
//   2  ROW part (and the row offsets) in shared memory, COLUMN part in a per-CTA area in global memory: a thread reads its
//      column's node once per rectangle, so the column data need not be close; without them two CTAs of 1 600-3 000-tip
//      pairs fit an SM where STG = 1 fits one (each thread then takes two or three columns of a rectangle in turn)
//   0  both parts in the per-CTA global area, read through L1: trees beyond the shared-memory limit (~3 500 tips).  The
//      reference has no size limit (phyloK2.py:158-209), and neither has this path.
template <typename T, bool CSMEM, int BOUND, bool DEAL, int STG>     // BOUND: upper bound of the CTA size (sets the register budget)
#ifndef PK_MINB192
#define PK_MINB192 5
#endif
#ifndef PK_MINB512
#define PK_MINB512 2
#endif
#ifndef PK_MINB192_DEAL
#define PK_MINB192_DEAL 4
#endif
#ifndef PK_MINB192_F32
#define PK_MINB192_F32 6
#endif
// resident CTAs per SM the register budget is set for.  <= 64 threads (100-tip trees): shared memory allows 12 (fp64,
// 80 registers) / 16 (fp32, 64 registers).  <= 192 threads: fp64 5 CTAs of 64 registers (pk_rect2 at 500 tips: 510 Gcells/s;
// 4 CTAs of 80 registers with 2 rows in flight 490, with 4 rows 515 at the price of spills); fp32 6 CTAs of 56 registers
// (902 against 881 with 5 and 814 with 4); labelled fp64 4 CTAs of 80 registers (3 tip states: shared memory allows no
// more; 2 states would run 5 % faster with 5), labelled fp32 5 of 64 (496 against 462 with 4).
__global__ void __launch_bounds__(BOUND, (STG == 2     ? 2
                                          : BOUND <= 64  ? (sizeof(T) == 4 ? 16 : 12)
                                          : BOUND <= 192 ? (sizeof(T) == 4 ? (DEAL ? 5 : PK_MINB192_F32) : DEAL ? PK_MINB192_DEAL : PK_MINB192)
                                          : BOUND <= 224 ? 4 : BOUND <= 256 ? PK_MINB : BOUND <= 352 ? 3 : BOUND <= 384 ? 2 : BOUND <= 512 ? PK_MINB512
                                          : BOUND <= 576 ? 2 : 1))
    pk_dp_kernel(const DpArgs a) {
    const int NT = blockDim.x, NW = NT >> 5;       // any multiple of 32 up to BOUND
    // rows in flight per lane: 2 where several light CTAs share an SM (more, lighter warps hide the latencies better:
    // measured 445 vs 410 Gcells/s at 500 tips) and wherever the register budget is tight (576-thread tier, 56 registers:
    // 432 vs 375 at 1500 tips), 4 for the big CTAs with room (1000 tips, 2 CTAs of 80 registers: 425 vs 397; 2000 tips:
    // 417 vs 363).  fp32 keeps 2 up to 576 threads (1300 tips: 807 vs 755 Gcells/s).
    constexpr int U = (sizeof(T) == 4 ? BOUND <= 576 : (BOUND <= PK_U2_BOUND || BOUND == 576)) ? PK_USMALL : PK_UNROLL;
    // where the unclamped prefetch pays (measured at 500 tips: fp32 744 -> 806 Gcells/s, labelled fp64 283 -> 297) and where
    // it does not (unlabelled fp64: ptxas then issues the gathers late, 452 -> 411)
    constexpr bool OVRQ = PK_OVERRUN != 0 && (PK_OVERRUN == 2 || sizeof(T) == 4 || DEAL);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ long long s_work[2];
    __shared__ int s_pair[2];
    __shared__ __align__(8) int s_state[4];      // pk_rect2: parity, first level of the batch, slab base
    __shared__ __align__(8) int s_cdist[2];      // labelled pk_rect2: distance of the b0 / b1 arrays, of the code arrays
    __shared__ int s_cntB[PK_MAX_CLASSES];
    __shared__ double s_part[32];
    __shared__ __align__(16) LevelDesc s_desc[PK_DESC_MAX];
    __shared__ __align__(16) unsigned long long s_tab[sizeof(T) == 8 ? PK_EXP_TAB * PK_EXP_REP + 4 : 2];   // + sigma, -, c3, c4 (pk_rect2)

    int tid;                                                   // opaque read: lives in a register, never re-read via S2R
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int lane = tid & 31, warp = tid >> 5;
    constexpr bool STAGED = STG != 0;                          // row records (and row offsets) in shared memory
    // child classes packed into the offset words (offsets < 2^24): where it measured faster -- fp64, unlabelled, the
    // 192-thread tier (200-550 tips: 468 against 460 Gcells/s at 500 tips, 348 against 342 at 200), the 320-thread tier
    // (800 tips: 442 against 425) and the 384-thread one (1000 tips: 434 against 432).  fp32 (788 against 805 at 500 tips),
    // the 64-thread tier, the big CTAs (2000 tips: 422 against 437) and the labelled variant lose with it.
    constexpr bool RECT2 = PK_RECT2 != 0 && !CSMEM && STG == 1 && (PK_RECT2 == 2 || (sizeof(T) == 8 && !DEAL));      // pk_rect2 (see there)
    constexpr int XO = RECT2 ? PK_XOFF : 0;                    // index of the first cell
    // descriptors with the column-data addresses, only the live rectangles visited, column words converted per pair (see the
    // wavefront): in the tiers with two rows in flight, except the 56-register fp64 tier of 576 threads (1500 tips: 490 against
    // 497 Gcells/s with it)
    constexpr bool D2 = PK_DESC2 != 0 && RECT2 && !DEAL && (PK_DESC2 == 2 || (U == 2 && !(sizeof(T) == 8 && BOUND == 576)));
    // labelled trees (chunk dealing): the same set-up, the two array distances in shared words of their own (the descriptor
    // needs its words for the chunk numbering)
    constexpr bool D2L = PK_DESC2L != 0 && RECT2 && DEAL && U == 2;
    constexpr bool CONVC = D2 || D2L;      // the column's words are converted when the pair is staged
    constexpr bool PQ = RECT2 || (PK_PACKQ != 0 && STG != 0 &&
                        (PK_PACKQ == 2 || (sizeof(T) == 8 && !DEAL && (BOUND == 192 || BOUND == 320 || BOUND == 384))));
    unsigned char *const gstage = STG == 1 ? nullptr : a.stage + (size_t)blockIdx.x * (size_t)a.stage_stride;
    unsigned char *sA = STAGED ? smem : gstage;                // ROW part of the row tree (its records are patched in place)
    unsigned char *sB = STG == 1 ? smem + a.tree_region : STG == 2 ? gstage : gstage + a.tree_region;   // COLUMN part of the column tree
    const size_t info_at = (size_t)a.tree_region + (STG == 2 ? 0 : a.col_region);
    int *rowoff = reinterpret_cast<int *>((STAGED ? smem : gstage) + info_at);
    const unsigned long long slab_off = (unsigned long long)blockIdx.x * (unsigned long long)a.slab_elems;
    Cells<T, CSMEM> cells;
    cells.sm = reinterpret_cast<T *>(smem + info_at + a.rowinfo_bytes);
    cells.base = (unsigned long long)(reinterpret_cast<T *>(a.scratch) + slab_off) - (unsigned long long)XO * sizeof(T);
#ifdef PK_DEBUG_BOUNDS
    cells.limit = (unsigned)a.slab_elems;
    cells.err = a.counter + 1;
    if (threadIdx.x == 0) {      // for pk_rect2 (visible behind the __syncthreads that follows the table fill)
        const unsigned long long lo = (unsigned long long)(reinterpret_cast<T *>(a.scratch) + slab_off);
        pk_dbg_lo[blockIdx.x & 4095u] = lo;
        pk_dbg_hi[blockIdx.x & 4095u] = lo + (unsigned long long)a.slab_elems * sizeof(T);
        pk_dbg_err[blockIdx.x & 4095u] = a.counter + 1;
    }
#endif
    if (tid == 0) mbar_init(&s_bar, 1);
    if (sizeof(T) == 8)
        for (int i = tid; i < PK_EXP_TAB * PK_EXP_REP; i += NT) s_tab[i] = a.exp_tab[i / PK_EXP_REP];
    if (sizeof(T) == 8 && tid == 0) {
        double *extra = reinterpret_cast<double *>(&s_tab[sizeof(T) == 8 ? PK_EXP_TAB * PK_EXP_REP : 0]);
        extra[0] = a.sigma;
        extra[2] = pk_poly[2];
        extra[3] = pk_poly[3];
    }
    if (tid < 32) s_part[tid] = 0.0;
    __syncthreads();
    // the slab base plus a zero read from a lane-dependent shared-memory address: ptxas cannot prove the sum uniform, so it stays in a vector
    // register pair and a cell address is one IMAD.WIDE idx, 8, base (a uniform base needs the 8 in a register: one
    // extra MOV per access)
    cells.base += (unsigned long long)*reinterpret_cast<volatile long long *>(&s_part[lane]);
    uint32_t parity = 0;
    const T lambda = (T)a.lambda, sigma = (T)a.sigma;
    const T tipf = sigma + lambda;                             // child factor tip x tip (phyloK2.py:194-196)
    GaussK<T> gk;
    gk.tab_addr = smem_u32(s_tab) + (lane & (PK_EXP_REP - 1)) * 8;
    gk.clamp_hi = a.clamp_hi;
    gk.lambda = lambda;
    gk.neg_inv_tau = (T)a.neg_inv_tau;
    gk.extra_addr = smem_u32(s_tab) + (sizeof(T) == 8 ? PK_EXP_TAB * PK_EXP_REP * 8u : 0u);
    gk.regc = false;
    gk.c3 = gk.c4 = 0.0;

    if (RECT2 && tid == 0) {      // pk_rect2 kernels keep per-CTA state in shared memory instead of registers (see the wavefront)
        s_state[0] = 0;           // mbarrier parity of the next staging
        s_state[1] = 0;           // first level of the current descriptor batch
        *reinterpret_cast<unsigned long long *>(&s_state[2]) = cells.base;
    }
    for (unsigned it_no = 0;; ++it_no) {
        const int slot = RECT2 ? 0 : (int)(it_no & 1);      // RECT2: one slot (every path to the next iteration passes a barrier after the read)
        if (tid == 0) s_work[slot] = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long w = s_work[slot];
        if (w >= a.total) break;

        // ---- which pair ------------------------------------------------------------------------
        int ia, ib;
        if (a.mode == 0) {
            const int2 p = a.pairs[w];
            ia = p.x;
            ib = p.y;
        } else if (a.mode == 1) {
            const int sh = a.tile_shift;
            const long long tile = w >> (2 * sh);
            const int r = (int)(w & ((1ll << (2 * sh)) - 1));
            const int2 t = a.tiles[tile];
            ia = (t.x << sh) + (r >> sh);
            ib = (t.y << sh) + (r & ((1 << sh) - 1));
            if (ib < ia || ib >= a.n_trees || ia >= a.n_trees) {
                if (RECT2) __syncthreads();      // everybody has read s_work[0] before thread 0 overwrites it
                continue;
            }
        } else {
            ia = ib = (int)w;
        }
        const int4 mA = a.metaA[2 * ia], mB = a.metaB[2 * ib];
        const bool swap = mA.w > mB.w;   // rows come from the tree with fewer wavefront levels (K is symmetric)
        const int4 mR = swap ? mB : mA;                                      // row part of the row tree
        const int4 mC = swap ? a.metaA[2 * ia + 1] : a.metaB[2 * ib + 1];    // column part of the column tree
        const unsigned char *srcR = (swap ? a.blocksB : a.blocksA) + (size_t)mR.x * 16;
        const unsigned char *srcC = (swap ? a.blocksA : a.blocksB) + (size_t)mC.x * 16;
        const uint32_t bytesR = mR.y, bytesC = mC.y;

        // ---- stage both node blocks with TMA bulk copies ---------------------------------------
        if (STAGED) {
            if (tid == 0) {
                // the staging buffers were last written through the generic proxy (record patch, length scale, pad records)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&s_bar, STG == 1 ? bytesR + bytesC : bytesR);
                tma_bulk_g2s(sA, srcR, bytesR, &s_bar);
                if (STG == 1) tma_bulk_g2s(sB, srcC, bytesC, &s_bar);
            }
        } else {
            // this CTA's private copies in global memory (the forest itself stays untouched: the records are patched per pair)
            for (uint32_t i = tid; i < bytesR / 16; i += NT) reinterpret_cast<int4 *>(sA)[i] = reinterpret_cast<const int4 *>(srcR)[i];
        }
        if (STG != 1) {
            for (uint32_t i = tid; i < bytesC / 16; i += NT) reinterpret_cast<int4 *>(sB)[i] = reinterpret_cast<const int4 *>(srcC)[i];
            __syncthreads();
        }
        if (STAGED) {
            if constexpr (RECT2) {
                mbar_wait(&s_bar, (uint32_t) * reinterpret_cast<volatile int *>(&s_state[0]));      // flipped by thread 0 behind the next barrier
            } else {
                mbar_wait(&s_bar, parity);
                parity ^= 1;
            }
        }
        if (OVRQ && !CSMEM && !RECT2 && tid < PK_PAD_RECORDS) {     // never-matching records after the last row (see PK_OVERRUN)
            int4 *padrec = reinterpret_cast<int4 *>(sA + bytesR) + 2 * tid;
            padrec[0] = make_int4(0, 0, 0, 0);
            padrec[1] = PQ ? make_int4((int)0xff000000u, (int)0xff000000u, 0, 0xffff) : make_int4(0, 0, 0, 0xffff);
        }

        const int *hA = reinterpret_cast<const int *>(sA);
        const int *hB = reinterpret_cast<const int *>(sB);
        const int nA = hA[0], ncls = hA[1], nlevA = hA[2], npc = 2 * ncls + 2;
        // classes whose rows all have a tip in slot 0 / slot 1 (those slots need no child-cell gathers)
        const unsigned long long tip0mask = (unsigned)hA[7] | ((unsigned long long)(unsigned)hA[8] << 32);
        const unsigned long long tip1mask = (unsigned)hA[9] | ((unsigned long long)(unsigned)hA[10] << 32);
        const int *lvlA = reinterpret_cast<const int *>(sA + hA[5]);
        RowRec<T> *rowrec = reinterpret_cast<RowRec<T> *>(sA + hA[6]);
        const int *clsB = reinterpret_cast<const int *>(sB + hB[4]);
        const int *gstB = reinterpret_cast<const int *>(sB + hB[5]);
        const T *b0p = reinterpret_cast<const T *>(sB + hB[6]);
        const T *b1p = reinterpret_cast<const T *>(sB + hB[7]);
        const int *codeB0 = reinterpret_cast<const int *>(sB + hB[8]);
        const int *codeB1 = reinterpret_cast<const int *>(sB + hB[9]);
        const int *pcB = reinterpret_cast<const int *>(sB + hB[10]);

        // ---- compact C layout.  Cells hold sigma + C (the factor a parent multiplies by).  Cell 0 = tip x tip, cell 1 = 1
        //      (child productions differ).  A cell (i, j) is only ever read by the cell of the two parents, and only when
        //      i and j have the same parent code; columns are ordered by (class, parent code), so row i stores just the
        //      contiguous column group with its own parent code: R cells per pair instead of M.
        if (tid < ncls) s_cntB[tid] = clsB[tid + 1] - clsB[tid];      // the host launches at least as many threads as classes
        if (tid == 0) {
            if constexpr (RECT2) {   // (indices start at XO: not through Cells::st, whose debug check counts from 0)
                stg_cell<T>(cells.base, XO, tipf);
                stg_cell<T>(cells.base, XO + 1, (T)1);
            } else {
                cells.st(XO, tipf);      // tip x tip child factor (phyloK2.py:194-196)
                cells.st(XO + 1, (T)1);  // factor of a child slot whose productions differ (phyloK2.py:191-192)
            }
        }
        // row offsets: exclusive scan of the stored row lengths, by warp 0 (32 rows per step, running carry)
        if (warp == 0) {
            int carry = 2 + XO;
            for (int base = 0; base < nA; base += 32) {
                const int r = base + lane;
                int len = 0;
                if (r < nA) {
                    const int pk = rowrec[r].packed;
                    const int g = ((pk >> 24) & 0xff) * npc + ((pk >> 16) & 0xff);
                    len = gstB[g + 1] - gstB[g];
                }
                int incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (r < nA) rowoff[r] = carry + incl - len;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        if (RECT2 && tid == 0) s_state[0] ^= 1;   // everybody is past this pair's mbarrier wait
        for (int r = tid; r < nA; r += NT) {      // patch the staged records with this pair's C offsets
            RowRec<T> *rec = rowrec + r;
            const int pk = rec->packed;
            const int c = (pk >> 24) & 0xff, p = (pk >> 16) & 0xff;
            const int o0 = (pk & 0xff) ? rowoff[rec->off0] : XO;
            const int o1 = (pk & 0xff00) ? rowoff[rec->off1] : XO;
            const int so = rowoff[r] - (gstB[c * npc + p] - clsB[c]);      // + column position in the class = stored cell
            rec->off0 = PQ ? (o0 | ((pk & 0xff) << 24)) : o0;              // PQ: the child's class + 1 in the top byte
            rec->off1 = PQ ? (o1 | ((pk & 0xff00) << 16)) : o1;
            rec->st_off = RECT2 ? (int)((unsigned)so | ((unsigned)p << 24)) : so;       // pk_rect2: the parent code rides in the top byte
            if constexpr (PQ && sizeof(T) == 4) {
                rec->st2 = RECT2 ? (int)((unsigned)so | ((unsigned)p << 24)) : so;
                rec->pk2 = pk;
            }
            rec->a0 *= (T)a.bscale;                // lengths in units that make the squared distance the exponent
            rec->a1 *= (T)a.bscale;
        }
        {
            T *b0w = const_cast<T *>(b0p), *b1w = const_cast<T *>(b1p);
            const int nB = hB[0];
            for (int c = tid; c < nB; c += NT) {
                b0w[c] *= (T)a.bscale;
                b1w[c] *= (T)a.bscale;
                if constexpr (CONVC) {   // the column's words as pk_rect2 uses them: rank - tag per child, parent code << 24
                    int *c0w = const_cast<int *>(codeB0), *c1w = const_cast<int *>(codeB1), *pw = const_cast<int *>(pcB);
                    const int x0 = c0w[c], x1 = c1w[c];
                    c0w[c] = (x0 & 0xffffff) - (int)((unsigned)x0 & 0xff000000u);
                    c1w[c] = (x1 & 0xffffff) - (int)((unsigned)x1 & 0xff000000u);
                    pw[c] = (int)((unsigned)pw[c] << 24);
                }
            }
        }

        // ---- wavefront over the row tree's levels, a batch of levels at a time -------------------------
        double acc = 0.0;
        if constexpr (RECT2) {
            // Same walk as the generic code below, written so that nothing but the running sum, the thread index and the loop
            // counters stays in registers across the row loops: whatever comes from the staged headers is read again where it
            // is needed (lds_i32 is volatile: not hoisted), the descriptors carry the shared-memory address of a rectangle's
            // first record, and the epilogue finds the pair in shared memory.  The row loops then keep their constants in
            // registers instead of re-reading them every iteration (64 registers per thread at 5 CTAs per SM).
            const uint32_t sA32 = smem_u32(sA), sB32 = smem_u32(sB);
            if (tid == 0) {
                s_pair[0] = ia;
                s_pair[1] = ib;
                s_state[1] = 0;
                if constexpr (D2L) {
                    s_cdist[0] = lds_i32(sB32 + 28) - lds_i32(sB32 + 24);
                    s_cdist[1] = lds_i32(sB32 + 36) - lds_i32(sB32 + 32);
                }
            }
            // one column of one rectangle (descriptor at shared-memory address DP): fields r0, rows, wB, colbase, chunk_end, flags, nchunks, address of the first record
#define PK_COLUMN2(DP, ROWS, JB)                                                                                                    \
    {                                                                                                                               \
        const uint32_t col = (uint32_t)(lds_i32((DP) + 12) + (JB));                                                                 \
        const T b0 = lds_real<T>(sB32 + lds_i32(sB32 + 24) + (uint32_t)sizeof(T) * col);                                            \
        const T b1 = lds_real<T>(sB32 + lds_i32(sB32 + 28) + (uint32_t)sizeof(T) * col);                                            \
        const int cb0 = lds_i32(sB32 + lds_i32(sB32 + 32) + 4u * col), cb1 = lds_i32(sB32 + lds_i32(sB32 + 36) + 4u * col);         \
        const int pc = lds_i32(sB32 + lds_i32(sB32 + 40) + 4u * col);                                                               \
        const uint32_t ra = (uint32_t)lds_i32((DP) + 28);                                                                           \
        unsigned long long base;                                                                                                    \
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(base) : "r"(smem_u32(&s_state[2])));                                          \
        switch (lds_i32((DP) + 20) & (512 | 1024)) {               /* bit 9 / 10: slot 0 / 1 is a tip in every row of the class */ \
        case 512 | 1024:                                                                                                            \
            acc = pk_rect2<T, false, false, U, DEAL, false>(ra, (ROWS), (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);            \
            break;                                                                                                                  \
        case 512:                                                                                                                   \
            acc = pk_rect2<T, false, true, U, DEAL, false>(ra, (ROWS), (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);             \
            break;                                                                                                                  \
        case 1024:                                                                                                                  \
            acc = pk_rect2<T, true, false, U, DEAL, false>(ra, (ROWS), (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);             \
            break;                                                                                                                  \
        default:                                                                                                                    \
            acc = pk_rect2<T, true, true, U, DEAL, false>(ra, (ROWS), (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);              \
        }                                                                                                                           \
    }
            // PK_DESC2 (unlabelled trees): the descriptor of a rectangle is {address of its first column's b0, rows, wB, address
            // of its first column's code0 | b1 - b0 array distance, flags, code array distance, address of the first record}, so a
            // thread sets up its column with two LDS.128 of the descriptor and five loads of its column's data -- eight loads
            // less per rectangle than going through the staged headers (~50 rectangles per 500-tip pair)
            // Only the live rectangles of a level get a descriptor (the level's last one says how many slots lie between it and
            // the next level's first), and the column's words are converted once per pair, when it is staged, into what pk_rect2
            // computes with (rank - tag per child, parent code << 24).
            // Measured per tier, fp64 Gcells/s: 192 threads (500 tips) 513 -> 529, 64 threads (100 tips) 229 -> 241, 200 tips
            // 388 -> 408, 320 threads 469 -> 478; fp32 500 tips 901 -> 969, 200 tips 611 -> 677, 1000 tips 862 -> 882.  The
            // tiers with four rows in flight and 80 registers lose with it (384 threads 511 -> 487, 768 threads 487 -> 469:
            // ptxas' schedule of the larger loop) and keep the walk through the staged headers.
#define PK_COLUMN2D(DP, Q1, JB)                                                                                                      \
    {                                                                                                                               \
        const int4 q2 = lds_i32x4((DP) + 16);                       /* {b1 - b0 distance, flags, code distance, first record} */    \
        PK_COLUMN2X(Q1, q2.x, q2.z, q2.y, q2.w, JB)                                                                                 \
    }
#define PK_COLUMN2X(Q1, DIST_B, DIST_C, FLAGS, RA, JB)                                                                              \
    {                                                                                                                               \
        const uint32_t ab = (uint32_t)(Q1).x + (uint32_t)sizeof(T) * (uint32_t)(JB), ac = (uint32_t)(Q1).w + 4u * (uint32_t)(JB);   \
        const T b0 = lds_real<T>(ab);                                                                                               \
        const T b1 = lds_real<T>(ab + (uint32_t)(DIST_B));                                                                          \
        const int cb0 = lds_i32(ac), cb1 = lds_i32(ac + (uint32_t)(DIST_C));                                                        \
        const int pc = lds_i32(ac + 2u * (uint32_t)(DIST_C));                                                                       \
        const uint32_t ra = (uint32_t)(RA);                                                                                         \
        unsigned long long base;                                                                                                    \
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(base) : "r"(smem_u32(&s_state[2])));                                          \
        switch ((FLAGS) & (512 | 1024)) {                                                                                           \
        case 512 | 1024:                                                                                                            \
            acc = pk_rect2<T, false, false, U, DEAL, true>(ra, (Q1).y, (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);            \
            break;                                                                                                                  \
        case 512:                                                                                                                   \
            acc = pk_rect2<T, false, true, U, DEAL, true>(ra, (Q1).y, (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);             \
            break;                                                                                                                  \
        case 1024:                                                                                                                  \
            acc = pk_rect2<T, true, false, U, DEAL, true>(ra, (Q1).y, (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);             \
            break;                                                                                                                  \
        default:                                                                                                                    \
            acc = pk_rect2<T, true, true, U, DEAL, true>(ra, (Q1).y, (JB), b0, b1, cb0, cb1, pc, base, gk, sigma, tipf, acc);              \
        }                                                                                                                           \
    }
            for (;;) {
                __syncthreads();                       // previous batch done with s_desc; patched records, s_state visible
                const int L0 = *reinterpret_cast<volatile int *>(&s_state[1]);
                if (L0 >= lds_i32(sA32 + 8)) break;
                const int ncls_a = lds_i32(sA32 + 4);
                const int lpb = DEAL ? max(1, PK_DESC_MAX / ncls_a) : PK_DESC_MAX / 3;      // levels per batch; unlabelled: 3 descriptor slots per level
                const int dstride = DEAL ? ncls_a : 3;
                {
                    const int nlev = lds_i32(sA32 + 8), L1 = min(nlev, L0 + lpb);
                    const uint32_t lvl32 = sA32 + lds_i32(sA32 + 20), rec32 = sA32 + lds_i32(sA32 + 24), cls32 = sB32 + lds_i32(sB32 + 16);
                    const unsigned long long tip0 = (unsigned)lds_i32(sA32 + 28) | ((unsigned long long)(unsigned)lds_i32(sA32 + 32) << 32);
                    const unsigned long long tip1 = (unsigned)lds_i32(sA32 + 36) | ((unsigned long long)(unsigned)lds_i32(sA32 + 40) << 32);
                    for (int L = L0 + tid; L < L1; L += NT) {          // one thread per level describes its rectangles
                        LevelDesc *dst = s_desc + (L - L0) * dstride;
                        int last = -1, chunks = 0;
                        int slot = 0;      // D2: the level's non-empty rectangles are written to its first slots
                        for (int c = 0; c < dstride; ++c) {
                            LevelDesc d;
                            d.r0 = 0; d.rows = 0; d.wB = 0; d.colbase = 0; d.flags = 0; d.nchunks = 0; d.chunk_end = 0; d.pad1 = 0;
                            if (c < ncls_a) {
                                const int r0 = lds_i32(lvl32 + 4 * (c * (nlev + 1) + L)), rows = lds_i32(lvl32 + 4 * (c * (nlev + 1) + L + 1)) - r0;
                                const int wB = s_cntB[c];
                                d.r0 = r0; d.wB = wB; d.colbase = lds_i32(cls32 + 4 * c);
                                d.pad1 = (int)(rec32 + 32u * (uint32_t)r0);          // shared-memory address of the first record
                                if constexpr (D2 || D2L) {      // (the code0, code1 and parent-code arrays are equally long)
                                    const int colbase = d.colbase;
                                    d.r0 = (int)(sB32 + (uint32_t)lds_i32(sB32 + 24) + (uint32_t)sizeof(T) * (uint32_t)colbase);
                                    d.colbase = (int)(sB32 + (uint32_t)lds_i32(sB32 + 32) + 4u * (uint32_t)colbase);
                                    if constexpr (D2) chunks = 0;
                                }
                                if (rows > 0 && wB > 0) {
                                    d.rows = rows;
                                    d.flags = (((tip0 >> c) & 1ull) ? 512 : 0) | (((tip1 >> c) & 1ull) ? 1024 : 0);
                                    last = c;
                                    chunks += (wB + 31) >> 5;
                                }
                            }
                            d.chunk_end = chunks;
                            if (!D2 && !DEAL && c == dstride - 1 && L == L1 - 1) d.flags |= 2048;      // last descriptor of the batch
                            if constexpr (D2) {
                                d.chunk_end = lds_i32(sB32 + 28) - lds_i32(sB32 + 24);     // b0 array -> b1 array
                                d.nchunks = lds_i32(sB32 + 36) - lds_i32(sB32 + 32);       // code0 -> code1 -> parent codes
                                // only the non-empty rectangles are visited: bits 12-13 of the flags = slots to the next
                                // descriptor (the level's last one leads to the next level's first slot)
                                if (d.rows > 0) {
                                    d.flags |= 1 << 12;
                                    dst[slot++] = d;
                                }
                            } else if constexpr (DEAL) {
                                // labelled trees: the live rectangles first, so that a warp looking for the rectangle of its
                                // chunk steps over no empty descriptors (1 + 2S + S^2 classes, half of them absent from a level)
                                if (d.rows > 0) dst[slot++] = d;
                            } else {
                                dst[c] = d;
                            }
                        }
                        if constexpr (DEAL) {
                            for (int k = slot; k < dstride; ++k) {
                                LevelDesc d;
                                d.r0 = 0; d.rows = 0; d.wB = 0; d.colbase = 0; d.flags = 0; d.nchunks = 0; d.chunk_end = chunks; d.pad1 = 0;
                                dst[k] = d;
                            }
                            if (L == L1 - 1) dst[dstride - 1].flags |= 2048;      // last descriptor of the batch
                            last = -1;
                        }
                        if constexpr (D2) {
                            if (slot == 0) {      // a level without live cells (the column tree lacks the classes of its rows)
                                LevelDesc d;
                                d.r0 = 0; d.rows = 0; d.wB = 0; d.colbase = 0; d.flags = 0; d.nchunks = 0; d.chunk_end = 0; d.pad1 = 0;
                                dst[slot++] = d;
                            } else {
                                dst[slot - 1].flags |= 256;
                            }
                            dst[slot - 1].flags = (dst[slot - 1].flags & ~(3 << 12)) | ((dstride - (slot - 1)) << 12) | (L == L1 - 1 ? 2048 : 0);
                        } else {
                            if (last >= 0) dst[last].flags |= 256;
                            dst[0].nchunks = chunks;
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) s_state[1] = L0 + lpb;       // read again behind the barrier at the loop's head
                if (DEAL) {
                    // the 32-column chunks of all rectangles of a level are numbered through and dealt to the warps round-robin
                    for (uint32_t dl = smem_u32(s_desc);; dl += (uint32_t)dstride * (uint32_t)sizeof(LevelDesc)) {
                        const int nchunks = lds_i32(dl + 24);
                        for (int ch = tid >> 5; ch < nchunks; ch += NT >> 5) {
                            uint32_t dp = dl;
                            while (ch >= lds_i32(dp + 16)) dp += (uint32_t)sizeof(LevelDesc);      // uniform per warp, at most ncls steps
                            if constexpr (D2L) {
                                const int4 q1 = lds_i32x4(dp), q2 = lds_i32x4(dp + 16);      // {b0 address, rows, wB, code0 address}, {chunk end, flags, -, first record}
                                const int2 dist = lds_i32x2(smem_u32(s_cdist));
                                const int jb = ((ch - (q2.x - ((q1.z + 31) >> 5))) << 5) + (tid & 31);
                                if (jb < q1.z) PK_COLUMN2X(q1, dist.x, dist.y, q2.y, q2.w, jb)
                                continue;
                            }
                            const int wB = lds_i32(dp + 8);
                            const int jb = ((ch - (lds_i32(dp + 16) - ((wB + 31) >> 5))) << 5) + (tid & 31);
                            if (jb < wB) PK_COLUMN2(dp, lds_i32(dp + 4), jb)
                        }
                        if (nchunks > 0) __syncthreads();    // end of a level: its cells are final before the next one reads
                        if (lds_i32(dl + (uint32_t)(dstride - 1) * (uint32_t)sizeof(LevelDesc) + 20) & 2048) break;
                    }
                } else if constexpr (D2) {
                    for (uint32_t dp = smem_u32(s_desc);;) {
                        const int4 q1 = lds_i32x4(dp);             // {b0 address, rows, wB, code0 address}
                        if (q1.y > 0) {                    // uniform
                            for (int jb = tid; jb < q1.z; jb += NT) PK_COLUMN2D(dp, q1, jb)
                        }
                        const int fl = lds_i32(dp + 20);
                        if (fl & 256) __syncthreads();      // end of a level: its cells are final before the next one reads
                        if (fl & 2048) break;
                        dp += (uint32_t)((fl >> 12) & 3) * (uint32_t)sizeof(LevelDesc);
                    }
                } else {
                    for (uint32_t dp = smem_u32(s_desc);; dp += (uint32_t)sizeof(LevelDesc)) {
                        const int rows = lds_i32(dp + 4);
                        if (rows > 0) {                    // uniform
                            // thread t owns column t; a rectangle wider than the CTA is finished by the first threads taking a second column
                            for (int jb = tid; jb < lds_i32(dp + 8); jb += NT) PK_COLUMN2(dp, rows, jb)
                        }
                        const int fl = lds_i32(dp + 20);
                        if (!(PK_WHATIF & 16) && rows > 0 && (fl & 256)) __syncthreads();      // end of a level: its cells are final before the next one reads
                        if (fl & 2048) break;
                    }
                }
            }
#undef PK_COLUMN2
#undef PK_COLUMN2D
#undef PK_COLUMN2X
        } else {
        const int lev_per_batch = max(1, PK_DESC_MAX / ncls);
        for (int L0 = 0; L0 < nlevA; L0 += lev_per_batch) {
            const int L1 = min(nlevA, L0 + lev_per_batch);
            __syncthreads();                       // previous batch done with s_desc; rowrec / s_coff visible
            for (int L = L0 + tid; L < L1; L += NT) {          // one thread per level describes its rectangles
                LevelDesc *dst = s_desc + (L - L0) * ncls;
                int last = -1, chunks = 0;
                for (int c = 0; c < ncls; ++c) {
                    const int r0 = lvlA[c * (nlevA + 1) + L], rows = lvlA[c * (nlevA + 1) + L + 1] - r0, wB = s_cntB[c];
                    LevelDesc d;
                    d.r0 = r0; d.rows = 0; d.wB = wB; d.colbase = clsB[c]; d.flags = 0; d.nchunks = 0; d.pad1 = 0;
                    if (rows > 0 && wB > 0) {
                        d.rows = rows;
                        d.flags = (((tip0mask >> c) & 1ull) ? 512 : 0) | (((tip1mask >> c) & 1ull) ? 1024 : 0);
                        last = c;
                        chunks += (wB + 31) >> 5;
                    }
                    d.chunk_end = chunks;
                    dst[c] = d;
                }
                if (last >= 0) dst[last].flags |= 256;
                dst[0].nchunks = chunks;
            }
            __syncthreads();
#define PK_COLUMN(D, JB)                                                                                              \
    {                                                                                                                 \
        const int col = (D).colbase + (JB);                                                                           \
        const T b0 = b0p[col], b1 = b1p[col];                                                                         \
        const int pkey = pcB[col] << 16;                  /* this column's parent code, aligned with rec.packed */   \
        const RowRec<T> *rp = rowrec + (D).r0;                                                                        \
        T part;                                                                                                       \
        switch ((D).flags & (512 | 1024)) {               /* bit 9 / 10: slot 0 / 1 is a tip in every row of the class */ \
        case 512 | 1024:                                                                                              \
            part = pk_rect<T, CSMEM, STAGED, false, false, U, DEAL || PK_FENCE_ALL, OVRQ, PQ>(rp, (D).rows, (JB), b0, b1, codeB0[col], codeB1[col], pkey,   \
                                                            cells, gk, sigma, tipf);                                  \
            break;                                                                                                    \
        case 512:                                                                                                     \
            part = pk_rect<T, CSMEM, STAGED, false, true, U, DEAL || PK_FENCE_ALL, OVRQ, PQ>(rp, (D).rows, (JB), b0, b1, codeB0[col], codeB1[col], pkey,    \
                                                           cells, gk, sigma, tipf);                                   \
            break;                                                                                                    \
        case 1024:                                                                                                    \
            part = pk_rect<T, CSMEM, STAGED, true, false, U, DEAL || PK_FENCE_ALL, OVRQ, PQ>(rp, (D).rows, (JB), b0, b1, codeB0[col], codeB1[col], pkey,    \
                                                           cells, gk, sigma, tipf);                                   \
            break;                                                                                                    \
        default:                                                                                                      \
            part = pk_rect<T, CSMEM, STAGED, true, true, U, DEAL || PK_FENCE_ALL, OVRQ, PQ>(rp, (D).rows, (JB), b0, b1, codeB0[col], codeB1[col], pkey,     \
                                                          cells, gk, sigma, tipf);                                    \
        }                                                                                                             \
        acc += (double)part;                                                                                          \
    }
            if (DEAL) {
                // The rectangles of a level are independent: their 32-column chunks are numbered through and dealt to the
                // warps round-robin, so a level of several narrow rectangles (1 + 2S + S^2 classes with S tip states)
                // keeps all warps busy instead of running the rectangles one after the other on a few warps each.
                for (int lv = 0; lv < L1 - L0; ++lv) {
                    const LevelDesc *dl = s_desc + lv * ncls;
                    const int nchunks = dl[0].nchunks;
                    for (int ch = warp; ch < nchunks; ch += NW) {
                        int e = 0;
                        while (ch >= dl[e].chunk_end) ++e;             // uniform per warp, at most ncls steps
                        const LevelDesc d = dl[e];
                        const int jb = ((ch - (d.chunk_end - ((d.wB + 31) >> 5))) << 5) + lane;
                        if (jb < d.wB) PK_COLUMN(d, jb)
                    }
                    if (nchunks > 0) __syncthreads();    // end of a level: its cells are final before the next one reads
                }
                continue;
            }
            const int ndesc = (L1 - L0) * ncls;
            for (int e = 0; e < ndesc; ++e) {
                const LevelDesc d = s_desc[e];
                if (d.rows <= 0) continue;             // uniform
                const int wB = d.wB;
                // thread t owns column t; a rectangle wider than the CTA (rare: the CTA is sized for the bulk of the
                // trees, not the widest) is finished by the first threads taking a second column
                for (int jb = tid; jb < wB; jb += NT) PK_COLUMN(d, jb)
                if (d.flags & 256) __syncthreads();      // end of a level: its cells are final before the next one reads
            }
        }
#undef PK_COLUMN
        }

        // ---- deterministic block reduction of K ---------------------------------------------------
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            if constexpr (RECT2) {      // kept in shared memory across the wavefront (see there)
                ia = *reinterpret_cast<volatile int *>(&s_pair[0]);
                ib = *reinterpret_cast<volatile int *>(&s_pair[1]);
            }
            double k = 0.0;
            for (int i = 0; i < NW; ++i) k += s_part[i];
            // every cell is a term of K, so a cell that overflowed (fp32 mode: beyond 3e38, which large decay factors reach
            // on deep similar trees) or an inf * 0 shows up here; the host entry points re-evaluate such pairs in fp64
            if (!(fabs(k) <= 1.7976931348623157e308)) atomicAdd(a.nonfinite, 1ull);
            if (a.mode == 0) {
                a.out[RECT2 ? *reinterpret_cast<volatile long long *>(&s_work[0]) : w] = k;
            } else if (a.mode == 1) {
                if (a.diag) k = k / sqrt(a.diag[ia] * a.diag[ib]);
                a.out[(long long)ia * a.ld + ib] = k;
                a.out[(long long)ib * a.ld + ia] = k;
            } else {
                a.out[ia] = k;
            }
        }
        // the next iteration's first __syncthreads orders these reads before the staging buffers are overwritten
    }
}

// decayFactor == 0: every C(n1, n2) is exactly 0 (phyloK2.py:184-199 multiplies by it), so K = 0 for every work item
// (0/0 = NaN where a normalisation by the zero diagonal is requested).  Same work decoding as pk_dp_kernel.
__global__ void pk_zero_kernel(const DpArgs a) {
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < a.total; w += (long long)gridDim.x * blockDim.x) {
        if (a.mode == 0) {
            a.out[w] = 0.0;
        } else if (a.mode == 1) {
            const int sh = a.tile_shift;
            const int2 t = a.tiles[w >> (2 * sh)];
            const int r = (int)(w & ((1ll << (2 * sh)) - 1));
            const int ia = (t.x << sh) + (r >> sh), ib = (t.y << sh) + (r & ((1 << sh) - 1));
            if (ib < ia || ib >= a.n_trees || ia >= a.n_trees) continue;
            const double k = a.diag ? 0.0 / sqrt(a.diag[ia] * a.diag[ib]) : 0.0;
            a.out[(long long)ia * a.ld + ib] = k;
            a.out[(long long)ib * a.ld + ia] = k;
        } else {
            a.out[w] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// context / forest
// ---------------------------------------------------------------------------------------------------
struct pk_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int cfg_threads = 0, cfg_ctas = 0, cfg_force_hbm = 0;
    // Packed image of the tree set that is uploaded again and again (a Gram matrix per step over the same trees): page-locked,
    // immutable, with the host-side statistics of its forest.  Built when a set of >= 64 trees is uploaded a second time;
    // from then on an upload of that set is one copy out of the image (no packing, no staging, no wait): packing 125 MB
    // for 4096 x 500-tip trees is bound by the host's memory bandwidth, whichever way it is spread over cores and ranks.
    std::vector<uint64_t> img_key, img_seen;    // tree uids + {precision, part_lo, part_hi}: of the image / of the last miss
    void *img_h = nullptr;
    size_t img_bytes = 0;
    pk_forest *img_forest = nullptr;            // host vectors only
    int opt_unstaged = 0;                       // 1: both tree parts in the global staging area even when they would fit (tests);
                                                // 2: COLUMN part in the global area wherever possible, 3: never
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    unsigned long long *d_counters = nullptr;   // ring of work counters (one per launch in flight)
    int counter_next = 0;
    void *d_tmp = nullptr;                      // staging for pair lists / tiles / results of the host-resident forms
    size_t tmp_bytes = 0;
    void *h_pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t pinned_ev = nullptr;            // recorded behind an async copy out of h_pinned that nobody waits for
    bool pinned_inflight = false;
    std::vector<std::pair<void *, size_t>> pool;   // device buffers of freed forests, reused by later uploads (all work
                                                   // is ordered on ctx->stream, so reuse needs no extra synchronisation)
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ev_valid = false;
    long long launches = 0;
    int last_path = 0;
    int last_grid = 0, last_threads = 0, last_smem = 0;
    std::map<std::tuple<int, int, int>, int> attr_smem;             // kernel variant -> dynamic shared-memory limit set
    std::map<std::tuple<int, int, int, int, int>, int> occupancy;   // (variant, threads, smem) -> resident CTAs per SM
};

struct pk_forest {
    pk_ctx *ctx = nullptr;
    size_t alloc_bytes = 0;
    int32_t n = 0;
    int32_t precision = 64;
    int32_t ncls = 3;
    unsigned char *d_blocks = nullptr;
    int4 *d_meta = nullptr;
    size_t bytes = 0;
    int32_t max_rowpart = 0, max_colpart = 0, max_n = 0;
    std::vector<int32_t> max_cnt;   // per class, max node count over the trees
    std::vector<int32_t> gcnt;      // [n][ncls * npc] nodes per (class, parent code) group
    std::vector<int32_t> max_gcnt;  // [ncls * npc] max over the trees
    std::vector<int32_t> tree_w;    // [n] widest class of each tree
    std::vector<int32_t> cnt;       // [n][ncls]
    std::vector<int64_t> hist;      // [n][ncls*2*(ncls+1)]
    std::vector<int32_t> nint;      // [n]
    std::vector<size_t> offs;       // [n + 1] byte offset of every tree's block in d_blocks
};

static const int kCounterRing = 64;

// Every entry point runs on its context's device and leaves the calling thread's current device as it found it (a
// torch caller with several GPUs does not expect a library call to switch devices under it).
struct DeviceGuard {
    int prev = -1, want;
    cudaError_t err;
    explicit DeviceGuard(int dev) : want(dev) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) {
            prev = -1;
            cudaGetLastError();
        }
        err = prev == dev ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
};
#define PK_ENTER(ctx)                                                                                       \
    DeviceGuard guard__((ctx)->device);                                                                     \
    if (guard__.err != cudaSuccess)                                                                         \
        return pk_fail(PK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(guard__.err));      \
    [[maybe_unused]] int rc = PK_OK

extern "C" int pk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pk_ctx_create(int device, pk_ctx **out) {
    if (!out) return pk_fail(PK_ERR_INVALID, "null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return pk_fail(PK_ERR_NODEVICE, std::string("no CUDA device available (") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                            "); this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return pk_fail(PK_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    PK_CUDA(cudaGetDeviceProperties(&prop, device));
    // the library carries one sm_100a cubin and no PTX: any other architecture (sm_103, sm_12x, ...) has no kernel image
    if (prop.major != 10 || prop.minor != 0)
        return pk_fail(PK_ERR_NODEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                            std::to_string(prop.minor) + "; this library is built for sm_100a only");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess)
        return pk_fail(PK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(guard.err));
    {
        cudaFuncAttributes fa;      // proves that the kernel image loads on this device before anything is allocated
        cudaError_t pe = cudaFuncGetAttributes(&fa, pk_zero_kernel);
        if (pe != cudaSuccess) {
            cudaGetLastError();
            return pk_fail(PK_ERR_NODEVICE, std::string("no kernel image for device '") + prop.name + "': " + cudaGetErrorString(pe));
        }
    }
    pk_ctx *ctx = new pk_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaError_t err = cudaSuccess;
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * (kCounterRing + 1));
    if (err == cudaSuccess) err = cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * (kCounterRing + 1));
    if (err == cudaSuccess) err = cudaEventCreate(&ctx->ev[0]);
    if (err == cudaSuccess) err = cudaEventCreate(&ctx->ev[1]);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&ctx->pinned_ev, cudaEventDisableTiming);
    if (err != cudaSuccess) {
        delete ctx;
        return pk_fail(PK_ERR_CUDA, std::string("pk_ctx_create: ") + cudaGetErrorString(err));
    }
    *out = ctx;
    return PK_OK;
}

extern "C" void pk_ctx_destroy(pk_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    for (auto &b : ctx->pool) cudaFree(b.first);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->img_h) cudaFreeHost(ctx->img_h);
    delete ctx->img_forest;
    if (ctx->ev[0]) cudaEventDestroy(ctx->ev[0]);
    if (ctx->ev[1]) cudaEventDestroy(ctx->ev[1]);
    if (ctx->pinned_ev) cudaEventDestroy(ctx->pinned_ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int pk_ctx_set_stream(pk_ctx *ctx, void *cuda_stream) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null ctx");
    PK_ENTER(ctx);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        PK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return PK_OK;
}

extern "C" int pk_ctx_sync(pk_ctx *ctx) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null ctx");
    PK_ENTER(ctx);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

extern "C" int pk_ctx_configure(pk_ctx *ctx, int32_t threads, int32_t ctas_per_sm, int32_t force_hbm) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null ctx");
    if (threads < 0 || threads > 1024 || (threads & 31))
        return pk_fail(PK_ERR_INVALID, "threads must be 0 or a multiple of 32 up to 1024");
    ctx->cfg_threads = threads;
    ctx->cfg_ctas = ctas_per_sm;
    ctx->cfg_force_hbm = force_hbm;
    return PK_OK;
}

extern "C" int pk_ctx_set_option(pk_ctx *ctx, const char *name, int32_t value) {
    if (!ctx || !name) return pk_fail(PK_ERR_INVALID, "null argument");
    const std::string key(name);
    if (key == "unstaged") {
        if (value < 0 || value > 3) return pk_fail(PK_ERR_INVALID, "unstaged: 0 (automatic), 1 (both parts global), 2 (column part global), 3 (never)");
        ctx->opt_unstaged = value;
    }
    else return pk_fail(PK_ERR_INVALID, "unknown option '" + key + "'");
    return PK_OK;
}

extern "C" int pk_ctx_last_stats(pk_ctx *ctx, double *ms, int64_t *launches, int32_t *path, int32_t launch_cfg[3]) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null ctx");
    PK_ENTER(ctx);
    if (ms) {
        *ms = 0.0;
        if (ctx->ev_valid) {
            PK_CUDA(cudaEventSynchronize(ctx->ev[1]));
            float f = 0.f;
            PK_CUDA(cudaEventElapsedTime(&f, ctx->ev[0], ctx->ev[1]));
            *ms = f;
        }
    }
    if (launches) *launches = ctx->launches;
    if (path) *path = ctx->last_path;
    if (launch_cfg) {
        launch_cfg[0] = ctx->last_grid;
        launch_cfg[1] = ctx->last_threads;
        launch_cfg[2] = ctx->last_smem;
    }
    return PK_OK;
}

extern "C" int pk_ctx_nonfinite(pk_ctx *ctx, int64_t *count, int reset) {
    if (!ctx || !count) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    unsigned long long v = 0;
    PK_CUDA(cudaMemcpyAsync(&v, ctx->d_counters + kCounterRing, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) PK_CUDA(cudaMemsetAsync(ctx->d_counters + kCounterRing, 0, sizeof(v), ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = (int64_t)v;
    return PK_OK;
}

static int ensure_tmp(pk_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->tmp_bytes) return PK_OK;
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    ctx->d_tmp = nullptr;
    ctx->tmp_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    PK_CUDA(cudaMalloc(&ctx->d_tmp, want));
    ctx->tmp_bytes = want;
    return PK_OK;
}

// called before every host write into the staging buffer
static int ensure_pinned(pk_ctx *ctx, size_t bytes) {
    if (ctx->pinned_inflight) {      // an earlier call left a copy out of the buffer queued (Gram tile list)
        PK_CUDA(cudaEventSynchronize(ctx->pinned_ev));
        ctx->pinned_inflight = false;
    }
    if (bytes <= ctx->pinned_bytes) return PK_OK;
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr;
    ctx->pinned_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    PK_CUDA(cudaMallocHost(&ctx->h_pinned, want));
    ctx->pinned_bytes = want;
    return PK_OK;
}

static int ensure_scratch(pk_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return PK_OK;
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
    PK_CUDA(cudaMalloc(&ctx->d_scratch, bytes));
    ctx->scratch_bytes = bytes;
    return PK_OK;
}

static cudaError_t pool_alloc(pk_ctx *ctx, size_t bytes, void **ptr, size_t *got) {
    int best = -1;
    for (size_t k = 0; k < ctx->pool.size(); ++k)
        if (ctx->pool[k].second >= bytes && ctx->pool[k].second <= 4 * bytes + (1 << 20) &&
            (best < 0 || ctx->pool[k].second < ctx->pool[best].second))
            best = (int)k;
    if (best >= 0) {
        *ptr = ctx->pool[best].first;
        *got = ctx->pool[best].second;
        ctx->pool.erase(ctx->pool.begin() + best);
        return cudaSuccess;
    }
    *got = std::max(bytes, (size_t)1 << 20);
    return cudaMalloc(ptr, *got);
}

static void pool_release(pk_ctx *ctx, void *ptr, size_t bytes) {
    ctx->pool.emplace_back(ptr, bytes);
    if (ctx->pool.size() > 6) {      // keep the pool small: drop the smallest buffer
        size_t k = 0;
        for (size_t i = 1; i < ctx->pool.size(); ++i)
            if (ctx->pool[i].second < ctx->pool[k].second) k = i;
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->pool[k].first);
        ctx->pool.erase(ctx->pool.begin() + k);
    }
}

// [part_lo, part_hi): the trees this call packs and copies to the device (all of them for a plain upload); the layout,
// the per-part meta records and the host-side statistics always cover the whole set
// tail_off != nullptr (transient forests of the host-tree entry points): the pinned staging buffer gets `tail_bytes` more
// after the forest (at *tail_off, for the caller's pair list and results) and the copy is NOT waited for -- the caller
// synchronises the stream once, after its own launch; until then it must not touch the staging buffer before *tail_off.
static int forest_upload_impl(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, int32_t precision, int32_t part_lo,
                              int32_t part_hi, pk_forest **out, size_t tail_bytes = 0, size_t *tail_off = nullptr) {
    if (!ctx || !trees || !out) return pk_fail(PK_ERR_INVALID, "null argument");
    if (n < 1) return pk_fail(PK_ERR_INVALID, "empty tree set");
    if (part_lo < 0 || part_hi > n || part_lo > part_hi) return pk_fail(PK_ERR_INVALID, "tree range out of bounds");
    if (precision != 64 && precision != 32) return pk_fail(PK_ERR_INVALID, "precision must be 64 or 32");
    *out = nullptr;
    PK_ENTER(ctx);
    static const bool trace = std::getenv("PK_TRACE") != nullptr;      // PK_TRACE=1: per-phase host times on stderr
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0.0;
    const int ncls = trees[0] ? trees[0]->ncls : 0;
    size_t total = 0;
    std::vector<uint32_t> rbs((size_t)n), cbs((size_t)n);      // part sizes, computed once per tree
    for (int32_t i = 0; i < n; ++i) {
        if (!trees[i]) return pk_fail(PK_ERR_INVALID, "null tree in set");
        if (trees[i]->ncls != ncls) return pk_fail(PK_ERR_INVALID, "trees use different production class schemes");
        rbs[i] = (uint32_t)pk_rowpart_bytes(*trees[i], precision);
        cbs[i] = (uint32_t)pk_colpart_bytes(*trees[i], precision);
        total += (size_t)rbs[i] + cbs[i];
    }
    if (total / 16 > 0x7fffffffull) return pk_fail(PK_ERR_INVALID, "tree set too large for one forest");
    const size_t meta_bytes = sizeof(int4) * 2 * (size_t)n;
    // ---- the set that is uploaded again and again: one copy out of its page-locked image -------------------------
    std::vector<uint64_t> key;
    const bool cacheable = !tail_off && n >= 64;
    if (cacheable) {
        key.resize((size_t)n + 3);
        for (int32_t i = 0; i < n; ++i) key[i] = trees[i]->uid;
        key[n] = (uint64_t)precision;
        key[(size_t)n + 1] = (uint64_t)part_lo;
        key[(size_t)n + 2] = (uint64_t)part_hi;
        if (ctx->img_forest && key == ctx->img_key) {
            pk_forest *f = new pk_forest(*ctx->img_forest);
            void *dptr = nullptr;
            cudaError_t e = pool_alloc(ctx, total + meta_bytes, &dptr, &f->alloc_bytes);
            f->d_blocks = static_cast<unsigned char *>(dptr);
            const unsigned char *img = static_cast<const unsigned char *>(ctx->img_h);
            if (e == cudaSuccess) {
                f->d_meta = reinterpret_cast<int4 *>(f->d_blocks + total);
                const size_t lo = f->offs[part_lo], hi = f->offs[part_hi];
                if (hi > lo) e = cudaMemcpyAsync(f->d_blocks + lo, img + lo, hi - lo, cudaMemcpyHostToDevice, ctx->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_meta, img + total, meta_bytes, cudaMemcpyHostToDevice, ctx->stream);
            }
            if (e != cudaSuccess) {
                if (f->d_blocks) cudaFree(f->d_blocks);
                delete f;
                cudaGetLastError();
                return pk_fail(PK_ERR_CUDA, std::string("pk_forest_upload: ") + cudaGetErrorString(e));
            }
            if (trace)
                std::fprintf(stderr, "pk_forest_upload: %d trees, %.1f MB out of the cached image: %.0f us\n", (int)(part_hi - part_lo),
                             1e-6 * (double)(f->offs[part_hi] - f->offs[part_lo]), 1e6 * (now() - t0));
            *out = f;
            return PK_OK;
        }
    }
    const int npc = 2 * ncls + 2;
    pk_forest *f = new pk_forest();
    f->ctx = ctx;
    f->n = n;
    f->precision = precision;
    f->ncls = ncls;
    f->bytes = total;
    f->max_cnt.assign(ncls, 0);
    f->cnt.resize((size_t)n * ncls);
    f->gcnt.resize((size_t)n * ncls * npc);
    f->max_gcnt.assign((size_t)ncls * npc, 0);
    f->hist.resize((size_t)n * ncls * 2 * (ncls + 1));
    f->nint.resize(n);
    f->tree_w.assign(n, 1);
    const size_t staged = (total + meta_bytes + 255) & ~(size_t)255;
    rc = ensure_pinned(ctx, staged + tail_bytes);
    if (rc) {
        delete f;
        return rc;
    }
    if (tail_off) *tail_off = staged;
    unsigned char *h = static_cast<unsigned char *>(ctx->h_pinned);
    int4 *hmeta = reinterpret_cast<int4 *>(h + total);
    std::vector<size_t> offs((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; ++i) offs[i + 1] = offs[i] + rbs[i] + cbs[i];
    // packing writes every byte of the forest once (125 MB for 4096 x 500 tips): spread it over the host cores
    auto pack_range = [&](int32_t lo, int32_t hi) {
        for (int32_t i = lo; i < hi; ++i) {
            const pk_tree &t = *trees[i];
            const size_t off = offs[i], rb = rbs[i];
            pk_rowpart_pack(t, precision, h + off);
            pk_colpart_pack(t, precision, h + off + rb);
        }
    };
    for (int32_t i = 0; i < n; ++i) {
        const pk_tree &t = *trees[i];
        const size_t off = offs[i], rb = rbs[i], cb = cbs[i];
        hmeta[2 * i] = make_int4((int)(off / 16), (int)rb, t.n, t.nlev);            // row part
        hmeta[2 * i + 1] = make_int4((int)((off + rb) / 16), (int)cb, t.n, t.nlev);   // column part
    }
    const double t1 = trace ? now() : 0.0;
    // Device buffer first, then pack and copy in phases: while the cores pack phase p + 1 into the pinned buffer, the
    // copy engine moves phase p (a 31 MB forest of 256 x 2000-tip trees: ~1.3 ms of copy hidden behind the packing).
    void *dptr = nullptr;
    cudaError_t e = pool_alloc(ctx, total + meta_bytes, &dptr, &f->alloc_bytes);
    f->d_blocks = static_cast<unsigned char *>(dptr);
    if (e == cudaSuccess) {
        f->d_meta = reinterpret_cast<int4 *>(f->d_blocks + total);
        const size_t part_bytes = offs[part_hi] - offs[part_lo];
        const int32_t pn = part_hi - part_lo;
        const int nthr = (int)std::min<size_t>(std::min(pk_host_threads(), 32), std::max<size_t>(1, part_bytes / ((size_t)128 << 10)));
        const int nphase = (int)std::min<size_t>(8, std::max<size_t>(1, part_bytes / ((size_t)4 << 20)));
        for (int ph = 0; ph < nphase && e == cudaSuccess; ++ph) {
            const int32_t p_lo = part_lo + (int32_t)((int64_t)pn * ph / nphase),
                          p_hi = part_lo + (int32_t)((int64_t)pn * (ph + 1) / nphase);
            std::atomic<int32_t> next{p_lo};
            const int32_t chunk = std::max<int32_t>(1, (p_hi - p_lo) / (nthr * 4));
            pk_parallel(nthr, [&](int) {
                for (;;) {
                    const int32_t lo = next.fetch_add(chunk);
                    if (lo >= p_hi) break;
                    pack_range(lo, std::min(p_hi, lo + chunk));
                }
            });
            if (offs[p_hi] > offs[p_lo])
                e = cudaMemcpyAsync(f->d_blocks + offs[p_lo], h + offs[p_lo], offs[p_hi] - offs[p_lo], cudaMemcpyHostToDevice,
                                    ctx->stream);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_meta, hmeta, meta_bytes, cudaMemcpyHostToDevice, ctx->stream);
    }
    const double t2 = trace ? now() : 0.0;
    // host-side statistics of every tree of the set (this rank's part or not): per-tree entries in parallel, maxima after
    {
        const int nthr = (int)std::min<size_t>(std::min(pk_host_threads(), 16), std::max<size_t>(1, (size_t)n / 256));
        std::atomic<int32_t> next{0};
        pk_parallel(nthr, [&](int) {
            for (;;) {
                const int32_t lo = next.fetch_add(64);
                if (lo >= n) break;
                for (int32_t i = lo; i < std::min(n, lo + 64); ++i) {
                    const pk_tree &t = *trees[i];
                    f->nint[i] = t.n;
                    int32_t wmax = 1;
                    for (int c = 0; c < ncls; ++c) {
                        const int cnt = t.cls_start[c + 1] - t.cls_start[c];
                        f->cnt[(size_t)i * ncls + c] = cnt;
                        wmax = std::max(wmax, cnt);
                    }
                    f->tree_w[i] = wmax;
                    for (int g = 0; g < ncls * npc; ++g) f->gcnt[(size_t)i * ncls * npc + g] = t.gstart[g + 1] - t.gstart[g];
                    std::copy(t.hist.begin(), t.hist.end(), f->hist.begin() + (size_t)i * t.hist.size());
                }
            }
        });
        for (int32_t i = 0; i < n; ++i) {
            f->max_rowpart = std::max<int32_t>(f->max_rowpart, (int32_t)rbs[i]);
            f->max_colpart = std::max<int32_t>(f->max_colpart, (int32_t)cbs[i]);
            f->max_n = std::max(f->max_n, f->nint[i]);
            for (int c = 0; c < ncls; ++c) f->max_cnt[c] = std::max(f->max_cnt[c], f->cnt[(size_t)i * ncls + c]);
            const int32_t *gc = &f->gcnt[(size_t)i * ncls * npc];
            for (int g = 0; g < ncls * npc; ++g) f->max_gcnt[g] = std::max(f->max_gcnt[g], gc[g]);
        }
    }
    const double t3 = trace ? now() : 0.0;
    if (e == cudaSuccess && !tail_off) e = cudaStreamSynchronize(ctx->stream);   // the pinned staging buffer is reused
    if (trace)
        std::fprintf(stderr, "pk_forest_upload: %d trees, %.1f MB: layout %.0f us, pack + queue copies %.0f us, stats %.0f us, wait %.0f us\n",
                     (int)(part_hi - part_lo), 1e-6 * (double)(offs[part_hi] - offs[part_lo]), 1e6 * (t1 - t0), 1e6 * (t2 - t1),
                     1e6 * (t3 - t2), 1e6 * (now() - t3));
    if (e != cudaSuccess) {
        if (f->d_blocks) cudaFree(f->d_blocks);
        delete f;
        cudaGetLastError();
        return pk_fail(PK_ERR_CUDA, std::string("pk_forest_upload: ") + cudaGetErrorString(e));
    }
    f->offs = offs;
    if (cacheable) {
        if (key == ctx->img_seen) {      // second upload of this set in a row: keep its image
            if (ctx->img_bytes < staged) {
                if (ctx->img_h) cudaFreeHost(ctx->img_h);
                ctx->img_h = nullptr;
                ctx->img_bytes = 0;
                if (cudaMallocHost(&ctx->img_h, staged) == cudaSuccess) ctx->img_bytes = staged;
                else cudaGetLastError();
            }
            delete ctx->img_forest;
            ctx->img_forest = nullptr;
            if (ctx->img_h) {
                // the staging buffer holds this rank's part of the blocks and the meta records (the copies out of it are done:
                // the stream was synchronised above)
                std::memcpy(static_cast<unsigned char *>(ctx->img_h) + offs[part_lo], h + offs[part_lo], offs[part_hi] - offs[part_lo]);
                std::memcpy(static_cast<unsigned char *>(ctx->img_h) + total, hmeta, meta_bytes);
                ctx->img_forest = new pk_forest(*f);
                ctx->img_forest->d_blocks = nullptr;
                ctx->img_forest->d_meta = nullptr;
                ctx->img_key = key;
            }
        }
        ctx->img_seen.swap(key);
    }
    *out = f;
    return PK_OK;
}

extern "C" int pk_forest_upload(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, int32_t precision, pk_forest **out) {
    return forest_upload_impl(ctx, trees, n, precision, 0, n, out);
}

extern "C" int pk_forest_upload_part(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, int32_t precision,
                                     int32_t part_lo, int32_t part_hi, pk_forest **out) {
    return forest_upload_impl(ctx, trees, n, precision, part_lo, part_hi, out);
}

extern "C" int pk_forest_device_range(const pk_forest *f, int32_t tree_lo, int32_t tree_hi, void **d_ptr, int64_t *bytes) {
    if (!f || !d_ptr || !bytes) return pk_fail(PK_ERR_INVALID, "null argument");
    if (tree_lo < 0 || tree_hi > f->n || tree_lo > tree_hi) return pk_fail(PK_ERR_INVALID, "tree range out of bounds");
    *d_ptr = f->d_blocks + f->offs[tree_lo];
    *bytes = (int64_t)(f->offs[tree_hi] - f->offs[tree_lo]);
    return PK_OK;
}

extern "C" void pk_forest_free(pk_forest *f) {
    if (!f) return;
    DeviceGuard guard(f->ctx->device);
    if (f->d_blocks) pool_release(f->ctx, f->d_blocks, f->alloc_bytes);
    delete f;
}

extern "C" int pk_forest_bytes(const pk_forest *f, int64_t *bytes) {
    if (!f || !bytes) return pk_fail(PK_ERR_INVALID, "null argument");
    *bytes = (int64_t)(f->bytes + sizeof(int4) * 2 * (size_t)f->n);
    return PK_OK;
}

// exact work counts {DPs, node pairs, matched cells M, child reads R}
static void pair_work(const pk_forest *fa, int i, const pk_forest *fb, int j, int64_t c[4]) {
    const int ncls = fa->ncls;
    const size_t hs = (size_t)ncls * 2 * (ncls + 1);
    int64_t m = 0, r = 0;
    for (int k = 0; k < ncls; ++k) {
        m += (int64_t)fa->cnt[(size_t)i * ncls + k] * fb->cnt[(size_t)j * ncls + k];
        for (int s = 0; s < 2; ++s)
            for (int q = 1; q <= ncls; ++q)
                r += fa->hist[(size_t)i * hs + ((size_t)k * 2 + s) * (ncls + 1) + q] *
                     fb->hist[(size_t)j * hs + ((size_t)k * 2 + s) * (ncls + 1) + q];
    }
    c[0] += 1;
    c[1] += (int64_t)fa->nint[i] * fb->nint[j];
    c[2] += m;
    c[3] += r;
}

extern "C" int pk_forest_pair_stats(const pk_forest *fa, const pk_forest *fb, const int32_t *pairs, int64_t npairs,
                                    int64_t counts[4]) {
    if (!fa || !fb || !pairs || !counts) return pk_fail(PK_ERR_INVALID, "null argument");
    if (fa->ncls != fb->ncls) return pk_fail(PK_ERR_INVALID, "forests use different production class schemes");
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        int i = pairs[2 * p], j = pairs[2 * p + 1];
        if (i < 0 || i >= fa->n || j < 0 || j >= fb->n) return pk_fail(PK_ERR_INVALID, "pair index out of range");
        pair_work(fa, i, fb, j, counts);
    }
    return PK_OK;
}

// ---------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------
// The dynamic shared-memory limit of a kernel variant is raised once per size and its occupancy is queried once per
// (CTA size, shared memory): a sim-vs-obs proposal is one 30 us launch, the driver calls would cost more than that.
template <typename T, bool CSMEM, int BOUND, bool DEAL, int STAGED>
static cudaError_t launch_t(pk_ctx *ctx, const DpArgs &args, int grid, int nt, int smem, cudaStream_t stream) {
    int &have = ctx->attr_smem[{(int)sizeof(T) + 16 * (int)DEAL + 32 * (int)STAGED, (int)CSMEM, BOUND}];
    if (smem > have) {
        cudaError_t e = cudaFuncSetAttribute(pk_dp_kernel<T, CSMEM, BOUND, DEAL, STAGED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    pk_dp_kernel<T, CSMEM, BOUND, DEAL, STAGED><<<grid, nt, smem, stream>>>(args);
    return cudaGetLastError();
}

template <typename T, bool CSMEM, int BOUND, bool DEAL, int STAGED>
static cudaError_t occupancy_t(pk_ctx *ctx, int nt, int smem, int *blocks) {
    auto key = std::make_tuple((int)sizeof(T) + 16 * (int)DEAL + 32 * (int)STAGED, (int)CSMEM, BOUND, nt, smem);
    auto it = ctx->occupancy.find(key);
    if (it != ctx->occupancy.end()) {
        *blocks = it->second;
        return cudaSuccess;
    }
    int &have = ctx->attr_smem[{(int)sizeof(T) + 16 * (int)DEAL + 32 * (int)STAGED, (int)CSMEM, BOUND}];
    if (smem > have) {
        cudaError_t e = cudaFuncSetAttribute(pk_dp_kernel<T, CSMEM, BOUND, DEAL, STAGED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, pk_dp_kernel<T, CSMEM, BOUND, DEAL, STAGED>, nt, smem);
    if (e == cudaSuccess) ctx->occupancy[key] = *blocks;
    return e;
}

// register budgets: <= 64 threads: 80 registers (fp32: 64), <= 192: 64 (5 CTAs per SM; labelled fp64 80), <= 224: 72
// (4 CTAs of 7 warps), <= 256: 80 (3 CTAs), <= 320: 64 and <= 352: 56 (3 CTAs; 800 tips: 432 Gcells/s against 390 in the 384 tier, 900 tips: 415 against 406),
// <= 384: 80 (2 CTAs), <= 576: 56 (2 CTAs; 1500 tips: 432 against 378 with one CTA), <= 512: 64 (2 CTAs; one CTA of 128 registers ran 5 %
// (fp64) / 37 % (fp32) slower at 1300 tips), <= 768: 80, else 64 (one CTA per SM)
#ifdef PK_DEV_TIER      // developer builds (scripts/dev_sass.sh): one kernel only (default: unlabelled fp64), compiles in seconds
#ifndef PK_DEV_T
#define PK_DEV_T double
#endif
#ifndef PK_DEV_DEAL
#define PK_DEV_DEAL false
#endif
#define PK_DISPATCH(FN, ...) FN<PK_DEV_T, false, PK_DEV_TIER, PK_DEV_DEAL, 1>(__VA_ARGS__)
#else
#define PK_DISPATCH_NT(FN, TT, CS, DL, ...)                                                     \
    (nt <= 64    ? FN<TT, CS, 64, DL, 1>(__VA_ARGS__)                                              \
     : nt <= 192 ? FN<TT, CS, 192, DL, 1>(__VA_ARGS__)                                           \
     : nt <= 224 ? FN<TT, CS, 224, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 256 ? FN<TT, CS, 256, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 320 ? FN<TT, CS, 320, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 352 ? FN<TT, CS, 352, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 384 ? FN<TT, CS, 384, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 512 ? FN<TT, CS, 512, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 576 ? FN<TT, CS, 576, DL, 1>(__VA_ARGS__)                                             \
     : nt <= 768 ? FN<TT, CS, 768, DL, 1>(__VA_ARGS__)                                             \
                 : FN<TT, CS, 1024, DL, 1>(__VA_ARGS__))
// unstaged: trees beyond the shared-memory staging limit (parts in a per-CTA global area, C in HBM slabs, 1024-thread tier)
// (colglobal: 1 600-3 000-tip trees, COLUMN part in the per-CTA global area, <= 384 threads taking several columns each)
#define PK_DISPATCH_CS(FN, TT, DL, ...)                          \
    (stg == 0   ? FN<TT, false, 1024, DL, 0>(__VA_ARGS__)        \
     : stg == 2 ? FN<TT, false, 384, false, 2>(__VA_ARGS__)      \
     : csmem    ? PK_DISPATCH_NT(FN, TT, true, DL, __VA_ARGS__)  \
                : PK_DISPATCH_NT(FN, TT, false, DL, __VA_ARGS__))
#define PK_DISPATCH(FN, ...)                                                                                \
    (prec32 ? (deal ? PK_DISPATCH_CS(FN, float, true, __VA_ARGS__) : PK_DISPATCH_CS(FN, float, false, __VA_ARGS__)) \
            : (deal ? PK_DISPATCH_CS(FN, double, true, __VA_ARGS__) : PK_DISPATCH_CS(FN, double, false, __VA_ARGS__)))
#endif

static int launch_dp(pk_ctx *ctx, const pk_forest *fa, const pk_forest *fb, DpArgs args, const pk_params *p) {
    if (fa->ctx != ctx || fb->ctx != ctx) return pk_fail(PK_ERR_INVALID, "forest belongs to another context");
    if (fa->precision != fb->precision) return pk_fail(PK_ERR_INVALID, "forests were uploaded with different precision");
    if (fa->ncls != fb->ncls) return pk_fail(PK_ERR_INVALID, "forests use different production class schemes");
    if (p->precision != 0 && p->precision != fa->precision)
        return pk_fail(PK_ERR_INVALID, "params.precision does not match the forest's precision");
    if (!(p->gauss != 0.0) || !std::isfinite(p->gauss) || !std::isfinite(p->decay) || !std::isfinite(p->sigma))
        return pk_fail(PK_ERR_INVALID, "decayFactor, gaussFactor, sigma must be finite and gaussFactor non-zero");
    if (args.total <= 0) return PK_OK;
    const bool prec32 = fa->precision == 32;
    const bool deal = fa->ncls > 3;      // labelled trees: many narrow class rectangles per level, dealt to the warps by chunks
    const size_t w = prec32 ? 4 : 8;

    args.blocksA = fa->d_blocks;
    args.metaA = fa->d_meta;
    args.blocksB = fb->d_blocks;
    args.metaB = fb->d_meta;
    args.lambda = p->decay;
    args.neg_inv_tau = -1.0 / p->gauss;
    args.sigma = p->sigma;
    if (!(p->gauss > 0.0)) return pk_fail(PK_ERR_INVALID, "gaussFactor must be positive");
    if (p->decay == 0.0) {
        pk_zero_kernel<<<(int)std::min<long long>((args.total + 255) / 256, 1024), 256, 0, ctx->stream>>>(args);
        PK_CUDA(cudaGetLastError());
        ctx->launches++;
        return PK_OK;
    }
    if (!(p->decay > 0.0) || p->decay < 1e-200 || p->decay > 1e200)
        return pk_fail(PK_ERR_INVALID, "decayFactor must be 0 or in [1e-200, 1e200]");
    if (!prec32) {
        // fp64 Gaussian (pk_gauss): table of lambda * 2^(j/64) with the high word biased by -(j << 14), the clamp that
        // keeps lambda * 2^(-E/64) a normal number, and the length scale that turns d^2/tau into 1/64 octaves
        for (int j = 0; j < 64; ++j) {
            const double t = p->decay * std::exp2((double)j / 64.0);
            unsigned long long bits;
            std::memcpy(&bits, &t, 8);
            const uint32_t hi = (uint32_t)(bits >> 32) - ((uint32_t)j << 14);
            args.exp_tab[j] = ((unsigned long long)hi << 32) | (bits & 0xffffffffull);
        }
        const int emin = 2 - 1023 - (int)std::floor(std::log2(p->decay));
        const double emax = -64.0 * (double)emin - 1.0;
        unsigned long long eb;
        std::memcpy(&eb, &emax, 8);
        args.clamp_hi = (int)(eb >> 32) - 1;
        args.bscale = std::sqrt(64.0 / (std::log(2.0) * p->gauss));
    } else {
        args.bscale = std::sqrt(1.0 / (std::log(2.0) * p->gauss));       // fp32: exponent in octaves
    }

    // Upper bound of the stored C cells over all pairs: row i keeps one cell per column of its own (class, parent code)
    // group, so a pair stores sum_g rows_g * cols_g cells; bound the columns by the per-group maximum over both forests.
    long long maxM = 0;
    {
        const int ng = fa->ncls * (2 * fa->ncls + 2);
        std::vector<long long> gmax(ng);
        for (int g = 0; g < ng; ++g) gmax[g] = std::max(fa->max_gcnt[g], fb->max_gcnt[g]);
        const pk_forest *fs[2] = {fa, fb};
        for (int k = 0; k < (fa == fb ? 1 : 2); ++k)
            for (int32_t i = 0; i < fs[k]->n; ++i) {
                long long m = 0;
                for (int g = 0; g < ng; ++g) m += (long long)fs[k]->gcnt[(size_t)i * ng + g] * gmax[g];
                maxM = std::max(maxM, m);
            }
        maxM += 2;   // cell 0 (tip x tip factor) and cell 1 (no-match factor) precede the rows
    }
    const int region_r = (std::max(fa->max_rowpart, fb->max_rowpart) + 32 * PK_PAD_RECORDS + 127) & ~127;
    const int region_c = (std::max(fa->max_colpart, fb->max_colpart) + 127) & ~127;
    const int max_n = std::max(fa->max_n, fb->max_n);
    const int rowinfo = ((max_n * 4) + 127) & ~127;                          // row offsets
    const long long base_smem = (long long)region_r + region_c + rowinfo;
    const long long limit = ctx->smem_optin - 16384;   // static shared memory (table, level descriptors, ...)
    // trees too large to stage in shared memory (~3 500 tips): their parts are staged in a per-CTA area in global memory
    int stg = (base_smem > limit || ctx->opt_unstaged == 1) ? 0 : 1;
    const long long c_bytes = ((maxM * (long long)w) + 127) & ~127ll;
    args.tree_region = region_r;
    args.col_region = region_c;
    args.rowinfo_bytes = rowinfo;

    // one thread per column of the widest class rectangle (column-stationary mapping)
    // CTA size: one thread per column for ~97 % of the trees (rounded up to a warp); the few wider rectangles are
    // finished by the first threads taking a second column, instead of every CTA carrying a mostly idle extra warp
    int max_w = 1;
    for (int c = 0; c < fa->ncls; ++c) max_w = std::max(max_w, std::max(fa->max_cnt[c], fb->max_cnt[c]));
    int typ_w;
    {
        std::vector<int32_t> ws(fa->tree_w);
        if (fb != fa) ws.insert(ws.end(), fb->tree_w.begin(), fb->tree_w.end());
        const size_t k = (size_t)std::min<double>((double)ws.size() - 1, std::floor(0.97 * (double)ws.size()));
        std::nth_element(ws.begin(), ws.begin() + k, ws.end());
        typ_w = ws[k];
    }
    int nt = std::min(1024, std::max(64, (typ_w + 31) & ~31));
    if (ctx->cfg_threads) nt = ctx->cfg_threads;
    if (nt < fa->ncls) nt = 64;      // the kernel fills its per-class tables with one thread per class (<= 64 classes)

    // C in shared memory only when that costs no resident CTAs: with C in a global slab the CTAs of small pairs stay
    // L2-resident anyway and more pairs per SM hide the level barriers better (measured: profiles/r01_summary.md)
    bool csmem = false;
    int smem = stg == 0 ? 0 : (int)base_smem, per_sm = 0;
    cudaError_t e = PK_DISPATCH(occupancy_t, ctx, nt, smem, &per_sm);
    // Big pairs (1 600 tips and more) fit one CTA per SM with both parts in shared memory.  With the COLUMN part in the
    // per-CTA global area two CTAs fit, each with half (a third, ...) as many threads, every thread taking that many
    // columns of a rectangle in turn: two independent pairs per SM overlap each other's level barriers and gather
    // latencies (opt_unstaged: 2 = wherever possible, 3 = never).
    if (e == cudaSuccess && stg == 1 && !deal && ctx->opt_unstaged != 3 && !ctx->cfg_threads &&
        (ctx->opt_unstaged == 2 || (per_sm == 1 && PK_COLGLOBAL_DEFAULT))) {
        int passes = 1;
        while (((((typ_w + passes - 1) / passes) + 31) & ~31) > 384) ++passes;
        if (ctx->opt_unstaged != 2 && passes == 1) passes = 2;
        const int nt2 = std::max(64, (((typ_w + passes - 1) / passes) + 31) & ~31);
        const int smem2 = region_r + rowinfo;
        int occ2 = 0;
        const int keep = stg;
        stg = 2;
        const int nt_keep = nt;
        nt = nt2;
        e = PK_DISPATCH(occupancy_t, ctx, nt, smem2, &occ2);
        if (e == cudaSuccess && smem2 <= limit && (occ2 >= 2 || ctx->opt_unstaged == 2) && occ2 >= 1) {
            smem = smem2;
            per_sm = occ2;
        } else {
            stg = keep;
            nt = nt_keep;
        }
    }
    if (e == cudaSuccess && stg == 1 && ctx->cfg_force_hbm != 1 && base_smem + c_bytes <= limit) {
        int occ_smem = 0;
        csmem = true;
        e = PK_DISPATCH(occupancy_t, ctx, nt, (int)(base_smem + c_bytes), &occ_smem);
        if (e == cudaSuccess && (occ_smem >= per_sm || ctx->cfg_force_hbm == 2)) {
            smem = (int)(base_smem + c_bytes);
            per_sm = occ_smem;
        } else {
            csmem = false;
        }
    }
    if (e != cudaSuccess) return pk_fail(PK_ERR_CUDA, std::string("occupancy query: ") + cudaGetErrorString(e));
    if (per_sm < 1) return pk_fail(PK_ERR_CUDA, "kernel does not fit on an SM with the requested configuration");
    if (!csmem && ctx->cfg_ctas > 0) per_sm = std::min(per_sm, ctx->cfg_ctas);
    long long grid_ll = std::min<long long>(args.total, (long long)per_sm * ctx->sm_count);
    int grid = (int)std::max(1ll, grid_ll);
    if (!csmem) {
        long long slack = 0;          // PK_OVERRUN: offset of any stored row + rank in any column group stays inside
        if (PK_OVERRUN)
            for (size_t g = 0; g < fa->max_gcnt.size(); ++g) slack = std::max<long long>(slack, std::max(fa->max_gcnt[g], fb->max_gcnt[g]));
        args.slab_elems = ((((maxM + slack) * (long long)w) + 255) & ~255ll) / (long long)w;
        if (PK_PACKQ && stg != 0 && args.slab_elems + PK_XOFF >= (1ll << 24))
            return pk_fail(PK_ERR_INVALID, "internal: staged pair with more than 2^24 stored cells");
        if (args.slab_elems >= (1ll << 31))
            return pk_fail(PK_ERR_INVALID, "tree pair too large: " + std::to_string(args.slab_elems) + " stored cells exceed the 32-bit cell index");
        // Every resident CTA owns one slab.  Huge pairs (tens of thousands of tips: hundreds of MB per slab) run on fewer
        // CTAs rather than asking for more memory than the device has.
        {
            const size_t per_cta = (size_t)args.slab_elems * w + (stg == 0 ? (size_t)base_smem + 256 : stg == 2 ? (size_t)region_c + 256 : 0);
            if ((size_t)grid * per_cta > ctx->scratch_bytes) {
                size_t free_b = 0, total_b = 0;
                if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                    const size_t budget = (size_t)(0.8 * (double)(free_b + ctx->scratch_bytes));
                    const long long fit = (long long)(budget / per_cta);
                    if (fit < 1)
                        return pk_fail(PK_ERR_NOMEM, "tree pair too large: one C slab of " + std::to_string(per_cta >> 20) + " MB does not fit the device memory");
                    grid = (int)std::min<long long>(grid, fit);
                } else {
                    cudaGetLastError();
                }
            }
        }
        // the slabs, then (unstaged) one staging area per CTA
        const size_t slab_bytes = ((size_t)grid * (size_t)args.slab_elems * w + 255) & ~(size_t)255;
        args.stage_stride = stg == 0 ? (long long)((base_smem + 255) & ~255ll) : stg == 2 ? (long long)((region_c + 255) & ~255) : 0;
        int rc = ensure_scratch(ctx, slab_bytes + (size_t)grid * (size_t)args.stage_stride);
        if (rc) return rc;
        args.scratch = ctx->d_scratch;
        args.stage = static_cast<unsigned char *>(ctx->d_scratch) + slab_bytes;
    }
#ifdef PK_DEBUG_BOUNDS
    if (csmem) args.slab_elems = c_bytes / (long long)w;                      // the limit the cell indices are checked against
    args.counter = ctx->d_counters + 2 * (ctx->counter_next++ % (kCounterRing / 2));   // [work counter, violations]
    PK_CUDA(cudaMemsetAsync(args.counter, 0, 2 * sizeof(unsigned long long), ctx->stream));
#else
    args.counter = ctx->d_counters + (ctx->counter_next++ % kCounterRing);
    PK_CUDA(cudaMemsetAsync(args.counter, 0, sizeof(unsigned long long), ctx->stream));
#endif
    args.nonfinite = ctx->d_counters + kCounterRing;
    PK_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    e = PK_DISPATCH(launch_t, ctx, args, grid, nt, smem, ctx->stream);
    if (e != cudaSuccess) return pk_fail(PK_ERR_CUDA, std::string("DP kernel launch: ") + cudaGetErrorString(e));
    PK_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
#ifdef PK_DEBUG_BOUNDS
    {
        unsigned long long bad = 0;
        PK_CUDA(cudaMemcpyAsync(&bad, args.counter + 1, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (bad) return pk_fail(PK_ERR_CUDA, "PK_DEBUG_BOUNDS: " + std::to_string(bad) + " cell accesses outside the C storage");
    }
#endif
    ctx->ev_valid = true;
    ctx->launches++;
    ctx->last_path = csmem ? 0 : (stg == 0 ? 2 : stg == 2 ? 3 : 1);
    ctx->last_grid = grid;
    ctx->last_threads = nt;
    ctx->last_smem = smem;
    return PK_OK;
}

extern "C" int pk_forest_kernel_pairs(pk_ctx *ctx, const pk_forest *fa, const pk_forest *fb, const int32_t *d_pairs,
                                      int32_t npairs, const pk_params *params, double *d_out_k) {
    if (!ctx || !fa || !fb || !params || (npairs > 0 && (!d_pairs || !d_out_k)))
        return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    DpArgs args;
    std::memset(&args, 0, sizeof(args));
    args.mode = 0;
    args.pairs = reinterpret_cast<const int2 *>(d_pairs);
    args.total = npairs;
    args.out = d_out_k;
    return launch_dp(ctx, fa, fb, args, params);
}

extern "C" int pk_forest_kernel_pairs_host(pk_ctx *ctx, const pk_forest *fa, const pk_forest *fb, const int32_t *pairs,
                                           int32_t npairs, const pk_params *params, double *out_k) {
    if (!ctx || !fa || !fb || !params || (npairs > 0 && (!pairs || !out_k))) return pk_fail(PK_ERR_INVALID, "null argument");
    if (npairs <= 0) return PK_OK;
    for (int32_t p = 0; p < npairs; ++p)
        if (pairs[2 * p] < 0 || pairs[2 * p] >= fa->n || pairs[2 * p + 1] < 0 || pairs[2 * p + 1] >= fb->n)
            return pk_fail(PK_ERR_INVALID, "pair index out of range");
    PK_ENTER(ctx);
    const size_t pb = sizeof(int32_t) * 2 * (size_t)npairs, ob = sizeof(double) * (size_t)npairs;
    const size_t pb_al = (pb + 255) & ~(size_t)255;
    rc = ensure_tmp(ctx, pb_al + ob);
    if (rc == PK_OK) rc = ensure_pinned(ctx, pb_al + ob);
    if (rc) return rc;
    unsigned char *d = static_cast<unsigned char *>(ctx->d_tmp);
    unsigned char *h = static_cast<unsigned char *>(ctx->h_pinned);
    std::memcpy(h, pairs, pb);
    PK_CUDA(cudaMemcpyAsync(d, h, pb, cudaMemcpyHostToDevice, ctx->stream));
    rc = pk_forest_kernel_pairs(ctx, fa, fb, reinterpret_cast<int32_t *>(d), npairs, params, reinterpret_cast<double *>(d + pb_al));
    if (rc) {
        std::string keep(pk_last_error());
        cudaStreamSynchronize(ctx->stream);
        pk_set_error(keep);
        return rc;
    }
    PK_CUDA(cudaMemcpyAsync(h + pb_al, d + pb_al, ob, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out_k, h + pb_al, ob);
    return PK_OK;
}

// `while_running`, if set, is called after the launch is queued and before the host waits for the results: host work
// that does not depend on them (kcoal) overlaps the kernel
static int kernel_pairs_impl(pk_ctx *ctx, const pk_tree *const *A, int32_t nA, const pk_tree *const *B, int32_t nB,
                             const int32_t *pairs, int32_t npairs, const pk_params *params, double *out_k,
                             const std::function<void()> *while_running) {
    if (!ctx || !A || !B || !params || (npairs > 0 && (!pairs || !out_k))) return pk_fail(PK_ERR_INVALID, "null argument");
    if (npairs <= 0) return PK_OK;
    for (int32_t p = 0; p < npairs; ++p)
        if (pairs[2 * p] < 0 || pairs[2 * p] >= nA || pairs[2 * p + 1] < 0 || pairs[2 * p + 1] >= nB)
            return pk_fail(PK_ERR_INVALID, "pair index out of range");
    PK_ENTER(ctx);
    const int prec = params->precision == 32 ? 32 : 64;
    pk_forest *fa = nullptr, *fb = nullptr;
    static const bool trace = std::getenv("PK_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0.0;
    const size_t pb = sizeof(int32_t) * 2 * (size_t)npairs, ob = sizeof(double) * (size_t)npairs;
    const size_t pb_al = (pb + 255) & ~(size_t)255;
    const bool same = (A == B && nA == nB);
    // one tree set (a proposal's replicates + the observed tree): its upload is not waited for, the pair list and the
    // results live behind it in the staging buffer and the stream is synchronised once, at the end
    size_t tail = 0;
    bool deferred = false;
    if (same) {
        rc = forest_upload_impl(ctx, A, nA, prec, 0, nA, &fa, pb_al + ob, &tail);
        deferred = rc == PK_OK;
    } else {
        rc = pk_forest_upload(ctx, A, nA, prec, &fa);
    }
    if (rc) return rc;
    if (trace) std::fprintf(stderr, "pk_kernel_pairs: forest upload %.0f us\n", 1e6 * (now() - t0));
    if (same) fb = fa;
    else rc = pk_forest_upload(ctx, B, nB, prec, &fb);
    if (rc == PK_OK) {
        rc = ensure_tmp(ctx, pb_al + ob);
        if (rc == PK_OK && !same) rc = ensure_pinned(ctx, pb_al + ob);
        if (rc == PK_OK) {
            unsigned char *d = static_cast<unsigned char *>(ctx->d_tmp);
            unsigned char *h = static_cast<unsigned char *>(ctx->h_pinned) + tail;
            std::memcpy(h, pairs, pb);
            cudaError_t e = cudaMemcpyAsync(d, h, pb, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) rc = pk_fail(PK_ERR_CUDA, cudaGetErrorString(e));
            pk_params pp = *params;
            pp.precision = prec;
            if (rc == PK_OK)
                rc = pk_forest_kernel_pairs(ctx, fa, fb, reinterpret_cast<int32_t *>(d), npairs, &pp,
                                            reinterpret_cast<double *>(d + pb_al));
            if (rc == PK_OK) {
                e = cudaMemcpyAsync(h + pb_al, d + pb_al, ob, cudaMemcpyDeviceToHost, ctx->stream);
                const double t1 = trace ? now() : 0.0;
                if (while_running) (*while_running)();
                const double t2 = trace ? now() : 0.0;
                if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
                if (trace)
                    std::fprintf(stderr, "pk_kernel_pairs: queued after %.0f us, host work while the kernel runs %.0f us, wait %.0f us\n",
                                 1e6 * (t1 - t0), 1e6 * (t2 - t1), 1e6 * (now() - t2));
                if (e != cudaSuccess) rc = pk_fail(PK_ERR_CUDA, std::string("pk_kernel_pairs: ") + cudaGetErrorString(e));
                else std::memcpy(out_k, h + pb_al, ob);
            }
        }
    }
    std::string keep = rc ? std::string(pk_last_error()) : std::string();
    if (rc && deferred) cudaStreamSynchronize(ctx->stream);      // failed before the final wait: the staging buffer may be in flight
    if (fb && fb != fa) pk_forest_free(fb);
    pk_forest_free(fa);
    if (rc) pk_set_error(keep);
    if (rc == PK_OK && prec == 32) {
        // fp32 cells overflow beyond 3e38 (decay factors above ~0.25 on deep similar trees, e.g. the reference's own unit-test
        // setting 0.5 on a caterpillar against itself): such a K arrives as inf / NaN and is evaluated again in fp64, which
        // has the reference's range -- no inf or NaN of the fp32 mode reaches the caller
        std::vector<int32_t> redo;
        std::vector<int32_t> where;
        for (int32_t p = 0; p < npairs; ++p)
            if (!std::isfinite(out_k[p])) {
                where.push_back(p);
                redo.push_back(pairs[2 * p]);
                redo.push_back(pairs[2 * p + 1]);
            }
        if (!where.empty()) {
            pk_params p64 = *params;
            p64.precision = 64;
            std::vector<double> k64(where.size());
            rc = kernel_pairs_impl(ctx, A, nA, B, nB, redo.data(), (int32_t)where.size(), &p64, k64.data(), nullptr);
            if (rc == PK_OK)
                for (size_t q = 0; q < where.size(); ++q) out_k[where[q]] = k64[q];
        }
    }
    return rc;
}

extern "C" int pk_kernel_pairs(pk_ctx *ctx, const pk_tree *const *A, int32_t nA, const pk_tree *const *B, int32_t nB,
                               const int32_t *pairs, int32_t npairs, const pk_params *params, double *out_k) {
    return kernel_pairs_impl(ctx, A, nA, B, nB, pairs, npairs, params, out_k, nullptr);
}

static int score_batch_impl(pk_ctx *ctx, const pk_tree *obs, const pk_tree *const *sims, int32_t nsims,
                            const pk_params *params, double *ref_denom, double *out_k, double *out_denom,
                            double *out_khat, const std::function<void()> *while_running) {
    if (!ctx || !obs || !sims || !params || !ref_denom) return pk_fail(PK_ERR_INVALID, "null argument");
    if (nsims < 1) return pk_fail(PK_ERR_INVALID, "no simulated trees");
    std::vector<const pk_tree *> all((size_t)nsims + 1);
    all[0] = obs;
    for (int32_t s = 0; s < nsims; ++s) all[s + 1] = sims[s];
    const bool need_ref = !(*ref_denom > 0.0);   // NaN or <= 0: compute K(obs, obs) (kamphir.py:157)
    const int32_t np = 2 * nsims + (need_ref ? 1 : 0);
    std::vector<int32_t> pairs(2 * (size_t)np);
    for (int32_t s = 0; s < nsims; ++s) {
        pairs[2 * (size_t)s] = 0;                       // k = kernel(target_tree, tree)      kamphir.py:290
        pairs[2 * (size_t)s + 1] = s + 1;
        pairs[2 * (size_t)(nsims + s)] = s + 1;         // tree_denom = kernel(tree, tree)    kamphir.py:291
        pairs[2 * (size_t)(nsims + s) + 1] = s + 1;
    }
    if (need_ref) pairs[2 * (size_t)(2 * nsims)] = pairs[2 * (size_t)(2 * nsims) + 1] = 0;
    std::vector<double> k((size_t)np);
    int rc = kernel_pairs_impl(ctx, all.data(), nsims + 1, all.data(), nsims + 1, pairs.data(), np, params, k.data(),
                               while_running);
    if (rc) return rc;
    if (need_ref) *ref_denom = k[2 * (size_t)nsims];
    for (int32_t s = 0; s < nsims; ++s) {
        if (out_k) out_k[s] = k[s];
        if (out_denom) out_denom[s] = k[(size_t)nsims + s];
        if (out_khat)                                   // kamphir.py:300, evaluated in log space like the reference
            out_khat[s] = std::exp(std::log(k[s]) - 0.5 * (std::log(*ref_denom) + std::log(k[(size_t)nsims + s])));
    }
    return PK_OK;
}

extern "C" int pk_score_batch(pk_ctx *ctx, const pk_tree *obs, const pk_tree *const *sims, int32_t nsims,
                              const pk_params *params, double *ref_denom, double *out_k, double *out_denom,
                              double *out_khat) {
    return score_batch_impl(ctx, obs, sims, nsims, params, ref_denom, out_k, out_denom, out_khat, nullptr);
}

extern "C" int pk_score_newick_batch(pk_ctx *ctx, const pk_tree *obs, const char *const *sims, const size_t *lens,
                                     int32_t nsims, int prep_flags, int normalize, int n_states, int nthreads,
                                     const pk_params *params, double *ref_denom, double kadj, double *out_k,
                                     double *out_denom, double *out_khat, double *out_kcoal, double *out_score,
                                     double *out_mean) {
    if (!ctx || !obs || !sims || !lens || !params || !ref_denom) return pk_fail(PK_ERR_INVALID, "null argument");
    if (nsims < 1) return pk_fail(PK_ERR_INVALID, "no simulated trees");
    pk_tree **trees = nullptr;
    int32_t n = 0;
    static const bool trace = std::getenv("PK_TRACE") != nullptr;      // PK_TRACE=1: per-phase host times on stderr
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0.0;
    const bool use_kcoal = !std::isnan(kadj);
    // kcoal is host work that does not depend on the DP (sorting a replicate's node heights, three dot products with the
    // target's): the worker that has just flattened a replicate computes it while the tree is in its cache, so it costs
    // no dispatch of its own (as a separate pass over the pool it was the longest host phase of a small proposal: 150 us
    // of 600 for 100 x 100-tip replicates).  A failure (a replicate with more internal nodes than the target, kamphir.py:264)
    // is reported after the DP, as before.
    std::vector<double> kc;
    int rc_kcoal = PK_OK;
    std::string err_kcoal;
    std::mutex kcoal_lock;
    std::function<void(size_t, pk_tree &)> post = [&](size_t k, pk_tree &t) {
        const int r = pk_kcoal(&t, obs, &kc[k]);
        if (r) {
            std::lock_guard<std::mutex> lk(kcoal_lock);
            if (rc_kcoal == PK_OK) {
                rc_kcoal = r;
                err_kcoal = std::string(pk_last_error());
            }
        }
    };
    // ... for small replicates.  Large ones keep the separate pass behind the launch: there the kernel runs long enough to
    // hide it (256 x 2000 tips: kernel 1.5 ms, sorting the heights 1.3 ms on 16 cores), and in front of the launch it would
    // not be hidden by anything.  The text length decides (24 KB is a ~700-tip tree; 100 x 500 tips: 75 us more in the
    // flattening pass against 120 us that stuck out behind the 100 us kernel).
    size_t text_bytes = 0;
    for (int32_t s = 0; s < nsims; ++s) text_bytes += lens[s];
    const bool kcoal_early = use_kcoal && text_bytes <= (size_t)nsims * 24576;
    if (use_kcoal) {
        kc.assign((size_t)nsims, std::numeric_limits<double>::quiet_NaN());
        if ((int32_t)obs->heights.size() == obs->n) pk_tree_sort_heights(obs);      // once, before the workers read them
    }
    int rc = pk_trees_from_newick_list_post(sims, lens, nsims, prep_flags, normalize, n_states, nthreads, &trees, &n,
                                            kcoal_early ? &post : nullptr);
    if (rc) return rc;
    const double t1 = trace ? now() : 0.0;
    std::vector<double> khat((size_t)n);
    // After the upload nothing reads the replicates' host trees any more; they are destroyed while the kernel runs, spread
    // over the pool (a tree is ~16 heap blocks: 116 us on one thread for 100 trees).
    auto free_trees = [&]() {
        if (!trees) return;
        const int workers = std::max(1, std::min(std::min(pk_host_threads(), 16), n / 8));
        pk_parallel(workers, [&](int wk) {
            for (int32_t s = wk; s < n; s += workers) pk_tree_free(trees[s]);
        });
        pk_tree_array_free(trees);
        trees = nullptr;
    };
    std::function<void()> overlap = [&]() {
        if (use_kcoal && !kcoal_early) {
            rc_kcoal = pk_kcoal_batch(trees, n, obs, kc.data());
            if (rc_kcoal) err_kcoal = pk_last_error();
        }
        free_trees();
    };
    rc = score_batch_impl(ctx, obs, trees, n, params, ref_denom, out_k, out_denom, khat.data(), &overlap);
    const double t2 = trace ? now() : 0.0;
    if (rc == PK_OK && use_kcoal && rc_kcoal) rc = pk_fail(rc_kcoal, (kcoal_early ? "replicate: " : "") + err_kcoal);
    if (rc == PK_OK) {
        double sum = 0.0;
        for (int32_t s = 0; s < n; ++s) {
            const double score = use_kcoal ? khat[s] * std::pow(kc[s], kadj) : khat[s];      // kamphir.py:306
            if (out_khat) out_khat[s] = khat[s];
            if (out_kcoal) out_kcoal[s] = use_kcoal ? kc[s] : std::numeric_limits<double>::quiet_NaN();
            if (out_score) out_score[s] = score;
            sum += score;
        }
        if (out_mean) *out_mean = sum / n;                                                   // kamphir.py:450
    }
    const double t3 = trace ? now() : 0.0;
    std::string keep = rc ? std::string(pk_last_error()) : std::string();
    free_trees();      // only if the call failed before the overlap window
    if (trace)
        std::fprintf(stderr, "pk_score_newick_batch: flatten %.0f us, score_batch %.0f us, kcoal+scores %.0f us, free %.0f us\n",
                     1e6 * (t1 - t0), 1e6 * (t2 - t1), 1e6 * (t3 - t2), 1e6 * (now() - t3));
    if (rc) pk_set_error(keep);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// Gram matrix
// ---------------------------------------------------------------------------------------------------
// Tiles of the upper triangle that `part` of `nparts` evaluates.  Without a forest: tile t (row-major over tj >= ti)
// belongs to part t % nparts.  With a forest: tiles are dealt by cost, cost(ti, tj) = sum over the tile's
// DPs of n_i * n_j (the node pairs, to which M and the DP time are proportional): longest-processing-time greedy --
// tiles by decreasing cost (ties in t order), each to the least loaded part (ties to the lowest part) -- then every part's
// list back in t order.  Every rank computes the same assignment from the same host-side statistics.  For a forest of
// equal-sized trees this deals the full tiles round-robin and then spreads the half-filled diagonal tiles evenly.
static int gram_tiles(int32_t n, int32_t tile, int32_t part, int32_t nparts, std::vector<int2> *tiles, int64_t *npairs,
                      const pk_forest *f = nullptr) {
    if (n < 1 || tile < 1 || (tile & (tile - 1)) || nparts < 1 || part < 0 || part >= nparts)
        return pk_fail(PK_ERR_INVALID, "gram: tile must be a power of two, 0 <= part < nparts");
    const int32_t nt = (n + tile - 1) / tile;
    auto tile_pairs = [&](int32_t ti, int32_t tj) {
        const int64_t hi = std::min<int64_t>(n, (int64_t)(ti + 1) * tile) - (int64_t)ti * tile;
        const int64_t wj = std::min<int64_t>(n, (int64_t)(tj + 1) * tile) - (int64_t)tj * tile;
        return (ti == tj) ? hi * (hi + 1) / 2 : hi * wj;
    };
    int64_t cnt = 0;
    if (!f || nparts == 1) {
        int64_t t = 0;
        for (int32_t ti = 0; ti < nt; ++ti)
            for (int32_t tj = ti; tj < nt; ++tj, ++t) {
                if (t % nparts != part) continue;
                if (tiles) tiles->push_back(make_int2(ti, tj));
                cnt += tile_pairs(ti, tj);
            }
        if (npairs) *npairs = cnt;
        return PK_OK;
    }
    std::vector<double> s1(nt, 0.0), s2(nt, 0.0);      // per block of trees: sum n, sum n^2
    for (int32_t i = 0; i < n; ++i) {
        s1[i / tile] += (double)f->nint[i];
        s2[i / tile] += (double)f->nint[i] * (double)f->nint[i];
    }
    struct Item { double cost; int64_t t; int32_t ti, tj; };
    std::vector<Item> items;
    items.reserve((size_t)nt * (nt + 1) / 2);
    int64_t t = 0;
    for (int32_t ti = 0; ti < nt; ++ti)
        for (int32_t tj = ti; tj < nt; ++tj, ++t)
            items.push_back(Item{ti == tj ? 0.5 * (s1[ti] * s1[ti] + s2[ti]) : s1[ti] * s1[tj], t, ti, tj});
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.cost > b.cost; });
    // least loaded part first; (load, part) pairs in a small binary heap
    std::vector<std::pair<double, int32_t>> heap;
    for (int32_t p = 0; p < nparts; ++p) heap.emplace_back(0.0, p);
    auto worse = [](const std::pair<double, int32_t> &a, const std::pair<double, int32_t> &b) { return a > b; };
    std::make_heap(heap.begin(), heap.end(), worse);
    std::vector<Item> mine;
    for (const Item &it : items) {
        std::pop_heap(heap.begin(), heap.end(), worse);
        if (heap.back().second == part) mine.push_back(it);
        heap.back().first += it.cost;
        std::push_heap(heap.begin(), heap.end(), worse);
    }
    std::sort(mine.begin(), mine.end(), [](const Item &a, const Item &b) { return a.t < b.t; });
    for (const Item &it : mine) {
        if (tiles) tiles->push_back(make_int2(it.ti, it.tj));
        cnt += tile_pairs(it.ti, it.tj);
    }
    if (npairs) *npairs = cnt;
    return PK_OK;
}

extern "C" int pk_gram_part_pairs(int32_t n, int32_t tile, int32_t part, int32_t nparts, int64_t *npairs) {
    if (!npairs) return pk_fail(PK_ERR_INVALID, "null argument");
    return gram_tiles(n, tile, part, nparts, nullptr, npairs);
}

extern "C" int pk_gram_part_tiles(int32_t n, int32_t tile, int32_t part, int32_t nparts, int32_t *tiles, int64_t *ntiles) {
    if (!ntiles) return pk_fail(PK_ERR_INVALID, "null argument");
    std::vector<int2> t;
    int rc = gram_tiles(n, tile, part, nparts, &t, nullptr);
    if (rc) return rc;
    if (tiles) {
        if (*ntiles < (int64_t)t.size()) return pk_fail(PK_ERR_INVALID, "tile buffer too small");
        for (size_t k = 0; k < t.size(); ++k) {
            tiles[2 * k] = t[k].x;
            tiles[2 * k + 1] = t[k].y;
        }
    }
    *ntiles = (int64_t)t.size();
    return PK_OK;
}

extern "C" int pk_forest_gram_part_tiles(const pk_forest *f, int32_t tile, int32_t part, int32_t nparts, int32_t *tiles,
                                         int64_t *ntiles) {
    if (!f || !ntiles) return pk_fail(PK_ERR_INVALID, "null argument");
    std::vector<int2> t;
    int rc = gram_tiles(f->n, tile, part, nparts, &t, nullptr, f);
    if (rc) return rc;
    if (tiles) {
        if (*ntiles < (int64_t)t.size()) return pk_fail(PK_ERR_INVALID, "tile buffer too small");
        for (size_t k = 0; k < t.size(); ++k) {
            tiles[2 * k] = t[k].x;
            tiles[2 * k + 1] = t[k].y;
        }
    }
    *ntiles = (int64_t)t.size();
    return PK_OK;
}

extern "C" int pk_forest_gram_stats(const pk_forest *f, int32_t tile, int32_t part, int32_t nparts, int64_t counts[4]) {
    if (!f || !counts) return pk_fail(PK_ERR_INVALID, "null argument");
    std::vector<int2> tiles;
    int rc = gram_tiles(f->n, tile, part, nparts, &tiles, nullptr, f);
    if (rc) return rc;
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (const int2 &t : tiles)
        for (int i = t.x * tile; i < std::min(f->n, (t.x + 1) * tile); ++i)
            for (int j = std::max(i, t.y * tile); j < std::min(f->n, (t.y + 1) * tile); ++j) pair_work(f, i, f, j, counts);
    return PK_OK;
}

extern "C" int pk_forest_gram_part(pk_ctx *ctx, const pk_forest *f, const pk_params *params, int normalised, int32_t tile,
                                   int32_t part, int32_t nparts, double *d_out, int64_t ld, double *d_diag) {
    if (!ctx || !f || !params || !d_out) return pk_fail(PK_ERR_INVALID, "null argument");
    if (normalised && !d_diag) return pk_fail(PK_ERR_INVALID, "normalised Gram needs a d_diag buffer");
    PK_ENTER(ctx);
    std::vector<int2> tiles;
    rc = gram_tiles(f->n, tile, part, nparts, &tiles, nullptr, f);
    if (rc) return rc;
    if (ld < f->n) return pk_fail(PK_ERR_INVALID, "gram: leading dimension smaller than N");
    if (d_diag) {   // K_ii for every tree (needed by every part for the normalisation)
        DpArgs dg;
        std::memset(&dg, 0, sizeof(dg));
        dg.mode = 2;
        dg.n_trees = f->n;
        dg.total = f->n;
        dg.out = d_diag;
        rc = launch_dp(ctx, f, f, dg, params);
        if (rc) return rc;
    }
    if (tiles.empty()) return PK_OK;
    const size_t tb = sizeof(int2) * tiles.size();
    rc = ensure_tmp(ctx, tb);
    if (rc == PK_OK) rc = ensure_pinned(ctx, tb);
    if (rc) return rc;
    std::memcpy(ctx->h_pinned, tiles.data(), tb);
    PK_CUDA(cudaMemcpyAsync(ctx->d_tmp, ctx->h_pinned, tb, cudaMemcpyHostToDevice, ctx->stream));
    PK_CUDA(cudaEventRecord(ctx->pinned_ev, ctx->stream));      // the launch is not waited for: the next writer of h_pinned is
    ctx->pinned_inflight = true;
    int shift = 0;
    while ((1 << shift) < tile) ++shift;
    DpArgs args;
    std::memset(&args, 0, sizeof(args));
    args.mode = 1;
    args.tiles = reinterpret_cast<const int2 *>(ctx->d_tmp);
    args.tile_shift = shift;
    args.n_trees = f->n;
    args.total = (long long)tiles.size() << (2 * shift);
    args.out = d_out;
    args.ld = ld;
    args.diag = normalised ? d_diag : nullptr;
    return launch_dp(ctx, f, f, args, params);
}

extern "C" int pk_gram(pk_ctx *ctx, const pk_tree *const *trees, int32_t n, const pk_params *params, int normalised,
                       double *out) {
    if (!ctx || !trees || !params || !out) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    const int prec = params->precision == 32 ? 32 : 64;
    pk_forest *f = nullptr;
    rc = pk_forest_upload(ctx, trees, n, prec, &f);
    if (rc) return rc;
    double *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, sizeof(double) * ((size_t)n * n + n));
    if (e != cudaSuccess) {
        pk_forest_free(f);
        return pk_fail(PK_ERR_CUDA, std::string("pk_gram: ") + cudaGetErrorString(e));
    }
    pk_params pp = *params;
    pp.precision = prec;
    const int32_t tile = n >= 1024 ? 32 : (n >= 128 ? 8 : 1);
    int64_t bad = 0;
    if (prec == 32) rc = pk_ctx_nonfinite(ctx, &bad, 1);      // start counting from zero
    if (rc == PK_OK) rc = pk_forest_gram_part(ctx, f, &pp, normalised, tile, 0, 1, d_out, n, d_out + (size_t)n * n);
    if (rc == PK_OK) {
        e = cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = pk_fail(PK_ERR_CUDA, std::string("pk_gram: ") + cudaGetErrorString(e));
    }
    if (rc == PK_OK && prec == 32) rc = pk_ctx_nonfinite(ctx, &bad, 1);
    std::string keep = rc ? std::string(pk_last_error()) : std::string();
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    pk_forest_free(f);
    if (rc) pk_set_error(keep);
    if (rc == PK_OK && prec == 32 && bad > 0) {      // fp32 cells overflowed somewhere: the whole matrix again in fp64 (rare)
        pp.precision = 64;
        return pk_gram(ctx, trees, n, &pp, normalised, out);
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// device buffers / IPC (one result matrix on rank 0 written by all ranks over NVLink)
// ---------------------------------------------------------------------------------------------------
extern "C" int pk_device_alloc(pk_ctx *ctx, size_t bytes, void **d_ptr) {
    if (!ctx || !d_ptr) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaMalloc(d_ptr, bytes));
    return PK_OK;
}
extern "C" int pk_device_free(pk_ctx *ctx, void *d_ptr) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaFree(d_ptr));
    return PK_OK;
}
/* page-locked host memory for results that are fetched every step (a 134 MB matrix lands in pinned memory at PCIe speed;
 * a fresh pageable array costs first-touch page faults and a staged copy) */
extern "C" int pk_host_alloc(pk_ctx *ctx, size_t bytes, void **h_ptr) {
    if (!ctx || !h_ptr) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaMallocHost(h_ptr, bytes));
    return PK_OK;
}
extern "C" int pk_host_free(pk_ctx *ctx, void *h_ptr) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaFreeHost(h_ptr));
    return PK_OK;
}
extern "C" int pk_ipc_export(pk_ctx *ctx, void *d_ptr, unsigned char handle[64]) {
    if (!ctx || !d_ptr || !handle) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PK_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, 64);
    return PK_OK;
}
extern "C" int pk_ipc_open(pk_ctx *ctx, const unsigned char handle[64], void **d_ptr) {
    if (!ctx || !d_ptr || !handle) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    PK_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PK_OK;
}
extern "C" int pk_ipc_close(pk_ctx *ctx, void *d_ptr) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return PK_OK;
}
extern "C" int pk_memcpy_d2h(pk_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
extern "C" int pk_memcpy_h2d(pk_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
extern "C" int pk_memset_d(pk_ctx *ctx, void *d_dst, int value, size_t bytes) {
    if (!ctx) return pk_fail(PK_ERR_INVALID, "null argument");
    PK_ENTER(ctx);
    PK_CUDA(cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    return PK_OK;
}
