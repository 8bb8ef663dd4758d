// Internal definitions shared by the host flattener (phylok_host.cpp) and the device side (phylok_dev.cu).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstddef>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/phylok.h"

#define PK_MAX_CLASSES 64          // 1 + 2S + S*S production classes with S <= 7 tip states
#define PK_HDR_INTS 16             // per-tree block header, 64 bytes

inline uint64_t pk_next_tree_uid() {
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}

// ---- one flattened tree (host) ---------------------------------------------------------------------
// Post-order arrays are exactly what annotate_tree (phyloK2.py:132-146) hangs on the Clade objects.
// "Kernel order" sorts the internal nodes by (production class, wavefront level, post-order index): with
// that order the production-matched node pairs of a tree pair are, per class, one dense cntA x cntB
// rectangle, and one wavefront level of the row tree is a contiguous row range of it.
struct pk_tree {
    int32_t n = 0;          // internal nodes
    int32_t ntips = 0;
    int32_t nlev = 0;       // wavefront levels (level 0 = nodes without internal children)
    int32_t ncls = 3;
    int32_t nstates = 0;
    double height = 0.0;    // un-normalised max depth
    double scale = 1.0;     // normalisation divisor
    std::vector<uint8_t> prod;      // [n]   production, 1..3
    std::vector<uint8_t> cls;       // [n]   class id in [0, ncls)  (unlabelled: prod - 1)
    std::vector<int32_t> cidx;      // [2n]  post-order index of internal child, -1 = tip
    std::vector<double> bl;         // [2n]  child branch lengths (after normalisation)
    std::vector<int32_t> level;     // [n]
    std::vector<int32_t> order;     // [n]   kernel row -> post-order index
    std::vector<int32_t> rank;      // [n]   post-order index -> rank within its class
    std::vector<int32_t> cls_start; // [ncls + 1]
    std::vector<int32_t> lvl;       // [ncls * (nlev + 1)]  absolute first kernel row of level L in class c
    std::vector<int64_t> hist;      // [ncls * 2 * (ncls + 1)]  nodes of class c whose slot-s child has class q (0 = tip)
    std::vector<double> heights;    // [n]   node heights, un-normalised; ascending once heights_sorted (first kcoal use)
    bool heights_sorted = false;
    std::once_flag heights_once;
    // column order: nodes sorted by (class, parent code, parent codes of the next ancestors, post-order).  The parent code of a node is
    // (class of its parent + 1) * 2 + its slot in the parent (0 for the root).  A C cell (i, j) is read -- by the cell of
    // the two parents -- only when i and j have the same parent code, so with this order the cells of row i that are
    // ever read are one contiguous column range, and only that range is stored.
    std::vector<uint8_t> pcode;     // [n]   post-order index -> parent code, < 2 * ncls + 2
    std::vector<int32_t> colorder;  // [n]   column -> post-order index
    std::vector<int32_t> krank;     // [n]   post-order index -> rank within its (class, parent code) group
    std::vector<int32_t> gstart;    // [ncls * npc + 1]  first column of group (class, parent code); npc = 2 * ncls + 2
    uint64_t uid = pk_next_tree_uid();   // never reused: identifies the tree in the per-context cache of packed tree sets
};

// Device block of one tree: two independently staged parts, both 16-byte aligned.
//   ROW part (the tree supplies the rows of C), kernel row order = (class, level, post-order):
//     int32 hdr[16] : 0 n | 1 ncls | 2 nlev | 3 part bytes | 4 off cls_start | 5 off lvl | 6 off row records |
//                     7,8 / 9,10 bitmask of the classes whose nodes all have a tip in slot 0 / slot 1
//     int32 cls_start[ncls + 1]; int32 lvl[ncls][nlev + 1];
//     record[n], 32 bytes each: real a0, a1; int32 child-0 row, child-1 row, 0, packed (q0 | q1 << 8 | parent code << 16
//     | class << 24, q = class + 1 of the child or 0 for a tip); the kernel patches the three ints per pair
//   COLUMN part (the tree supplies the columns), column order = (class, parent code, post-order):
//     int32 hdr[16] : 0 n | 1 ncls | 3 part bytes | 4 off cls_start | 5 off gstart | 6 off bl0 | 7 off bl1 | 8 off code0 |
//                     9 off code1 | 10 off pcode
//     int32 cls_start[ncls + 1]; int32 gstart[ncls * npc + 1]; real bl0[n], bl1[n];
//     int32 code0[n], code1[n]   child: (class + 1) << 24 | rank of the child within its (class, parent code) group
//     int32 pcode[n]
size_t pk_rowpart_bytes(const pk_tree &t, int precision);
size_t pk_colpart_bytes(const pk_tree &t, int precision);
void pk_rowpart_pack(const pk_tree &t, int precision, unsigned char *dst);
void pk_colpart_pack(const pk_tree &t, int precision, unsigned char *dst);

// runs fn(worker) on `workers` threads (the caller is worker 0, the rest come from a process-wide pool of parked
// threads: starting a std::thread per call costs more than flattening a small batch of trees).  Fork-safe: a child
// process builds its own pool on first use.
void pk_parallel(int workers, const std::function<void(int)> &fn);
int pk_host_threads();

void pk_set_error(const std::string &msg);
int pk_fail(int code, const std::string &msg);

// pk_trees_from_newick_list with a follow-up per tree, run by the worker thread that flattened it (tree index, tree)
int pk_trees_from_newick_list_post(const char *const *texts, const size_t *lens, int32_t n, int prep_flags, int normalize,
                                   int n_states, int nthreads, pk_tree ***out, int32_t *n_out,
                                   const std::function<void(size_t, pk_tree &)> *post);
// sorts the node heights of a tree (kcoal's first step) now instead of on first use
void pk_tree_sort_heights(const pk_tree *t);

// finishes a tree whose post-order arrays (prod/cls/cidx/bl, n, ncls) are filled: levels, kernel order, histograms
int pk_tree_finish(pk_tree &t, bool levels_known = false);
