// Host flattener: Newick text -> ladderized, rescaled, annotated, kernel-ordered SoA tree.
//
// Replaces, for the kernel path only, what the reference does with Bio.Phylo objects:
//   Phylo.read / Phylo.parse            kamphir.py:112,323,373
//   root.branch_length = 0; ladderize   kamphir.py:153-154, 279-280
//   PhyloKernel.normalize_tree          phyloK2.py:101-130
//   PhyloKernel.annotate_tree           phyloK2.py:132-146
//   node heights for kcoal              kamphir.py:255-258
// Structural parity with the reference (child order after the stable ladderize, post-order numbering,
// pre-order summation order of total_branch_length) is tested bit-for-bit against the oracle in
// tests/test_flatten.py.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <vector>

#include "phylok_internal.h"

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

void pk_set_error(const std::string &msg) { g_err = msg; }
int pk_fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

extern "C" const char *pk_last_error(void) { return g_err.c_str(); }
extern "C" int pk_abi_version(void) { return PK_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------------
// worker pool
// ---------------------------------------------------------------------------------------------------
namespace {
struct WorkPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    const std::function<void(int)> *job = nullptr;
    unsigned long generation = 0;
    int wanted = 0;      // pool threads that take part in the current job (ids 1..wanted)
    int pending = 0;
    std::mutex run_lock; // one job at a time

    explicit WorkPool(int n) {
        for (int i = 0; i < n; ++i) threads.emplace_back([this, i]() { loop(i + 1); });
    }
    void loop(int id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)> *fn;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&]() { return generation != seen; });
                seen = generation;
                if (id > wanted) continue;
                fn = job;
            }
            (*fn)(id);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    void run(int workers, const std::function<void(int)> &fn) {
        std::lock_guard<std::mutex> serial(run_lock);
        const int extra = std::min<int>(workers - 1, (int)threads.size());
        if (extra > 0) {
            std::lock_guard<std::mutex> lk(m);
            job = &fn;
            wanted = extra;
            pending = extra;
            ++generation;
        }
        if (extra > 0) cv_work.notify_all();
        fn(0);
        if (extra > 0) {
            std::unique_lock<std::mutex> lk(m);
            cv_done.wait(lk, [&]() { return pending == 0; });
        }
    }
};
std::atomic<WorkPool *> g_pool{nullptr};
std::mutex g_pool_make;
std::once_flag g_atfork_once;
}  // namespace

// Host threads this process may use: the hardware concurrency, divided by the number of ranks that share the host when
// a launcher says so (torchrun sets LOCAL_WORLD_SIZE: 8 ranks each packing their part of a forest on 32 threads of a
// 32-core host only get in each other's way); PK_HOST_THREADS overrides.
int pk_host_threads() {
    static const int n = []() {
        int hc = (int)std::max(1u, std::thread::hardware_concurrency());
        if (const char *e = std::getenv("PK_HOST_THREADS")) {
            const int v = std::atoi(e);
            if (v > 0) return v;
        }
        if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) {
            const int v = std::atoi(e);
            if (v > 1) hc = std::max(1, hc / v);
        }
        return hc;
    }();
    return n;
}

void pk_parallel(int workers, const std::function<void(int)> &fn) {
    if (workers <= 1) {
        fn(0);
        return;
    }
    WorkPool *p = g_pool.load(std::memory_order_acquire);
    if (!p) {
        std::lock_guard<std::mutex> lk(g_pool_make);
        p = g_pool.load(std::memory_order_acquire);
        if (!p) {
            // after fork() the parent's pool threads do not exist in the child: forget the pool (leaked on purpose,
            // its mutexes may be in any state) and let the child build its own
            std::call_once(g_atfork_once, []() { pthread_atfork(nullptr, nullptr, []() { g_pool.store(nullptr); }); });
            p = new WorkPool(std::min(pk_host_threads(), 64) - 1);
            g_pool.store(p, std::memory_order_release);
        }
    }
    p->run(workers, fn);
}

// ---------------------------------------------------------------------------------------------------
// Newick
// ---------------------------------------------------------------------------------------------------
namespace {

// one node per token-created clade; all fields of a node share a cache line (the parser appends one node per '(' or ',')
struct RawNode {
    int32_t parent, first, last, next, nchild, name_off, name_len;
    int32_t ntip;                            // terminals below (the node itself if it is one); complete once the node is closed
    double bl;                               // NaN = absent (Biopython: None)
};
struct RawTree {
    std::vector<RawNode> nd;
    int32_t root = -1;

    int32_t add(int32_t par) {
        const int32_t id = (int32_t)nd.size();
        nd.push_back(RawNode{par, -1, -1, -1, 0, 0, 0, 0, std::numeric_limits<double>::quiet_NaN()});
        if (par >= 0) link(par, id);
        return id;
    }
    void link(int32_t par, int32_t id) {
        nd[id].parent = par;
        nd[id].next = -1;
        RawNode &p = nd[par];
        if (p.first < 0) p.first = id; else nd[p.last].next = id;
        p.last = id;
        p.nchild++;
    }
    size_t size() const { return nd.size(); }
    void clear() {      // keeps the capacity: one RawTree is reused for every tree a worker parses
        nd.clear();
        root = -1;
    }
    void reserve(size_t n) { nd.reserve(n); }
};

// per-worker scratch of flatten(): capacities survive from tree to tree
struct Scratch {
    RawTree raw;
    std::vector<int32_t> kids, pidx, lvl;
    std::vector<double> depth, lens;
    std::vector<std::pair<int32_t, int32_t>> st;
    std::string err;
};

// character classes of the Newick tokenizer: 1 = punctuation, 2 = white space
struct CharClass {
    unsigned char t[256];
    constexpr CharClass() : t() {
        for (int i = 0; i < 256; ++i) t[i] = 0;
        for (char c : {'(', ')', '[', ']', '\'', ':', ';', ','}) t[(unsigned char)c] = 1;
        for (char c : {' ', '\t', '\n', '\r', '\f', '\v'}) t[(unsigned char)c] = 2;
    }
};
constexpr CharClass kCls;
inline bool is_punct(char c) { return kCls.t[(unsigned char)c] == 1; }
inline bool is_space(char c) { return kCls.t[(unsigned char)c] == 2; }
inline bool is_token_end(char c) { return kCls.t[(unsigned char)c] != 0; }

// eight ASCII digits at once (little-endian word, first character in the low byte)
inline bool all_digits8(uint64_t w) {
    return (((w & 0xF0F0F0F0F0F0F0F0ull) | (((w + 0x0606060606060606ull) & 0xF0F0F0F0F0F0F0F0ull) >> 4)) ==
            0x3333333333333333ull);
}
inline uint32_t digits8(uint64_t w) {
    w = (w & 0x0F0F0F0F0F0F0F0Full) * 2561 >> 8;                  // pairs: d0 * 10 + d1
    w = (w & 0x00FF00FF00FF00FFull) * 6553601 >> 16;              // quads
    return (uint32_t)((w & 0x0000FFFF0000FFFFull) * 42949672960001ull >> 32);
}

// Branch lengths: digits[.digits][e[+-]digits] with at most 19 significant digits and a power of ten within 10^22 is
// the exact case of Clinger's fast path -- the decimal mantissa fits a double exactly (< 2^53) and so does the power
// of ten, so one correctly rounded multiplication or division gives the correctly rounded result, the same double
// strtod / Python's float() produce.  Everything else (longer mantissas, huge exponents, inf, nan, hex) goes to
// std::from_chars.
inline bool parse_length(const char *p, const char *end, double &out) {
    static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char *q = p;
    bool neg = false;
    if (q < end && *q == '-') { neg = true; ++q; }
    uint64_t m = 0;
    int digits = 0, frac = 0;
    bool any = false;
    while (q < end && (unsigned)(*q - '0') < 10u) {
        if (m || *q != '0') { if (++digits > 19) goto slow; m = m * 10 + (unsigned)(*q - '0'); }
        any = true;
        ++q;
    }
    if (q < end && *q == '.') {
        ++q;
        while (q < end && (unsigned)(*q - '0') < 10u) {
            if (m || *q != '0') { if (++digits > 19) goto slow; m = m * 10 + (unsigned)(*q - '0'); }
            any = true;
            ++frac;
            ++q;
        }
    }
    if (!any) goto slow;
    {
        int ex = 0;
        if (q < end && (*q == 'e' || *q == 'E')) {
            ++q;
            bool eneg = false;
            if (q < end && (*q == '+' || *q == '-')) { eneg = *q == '-'; ++q; }
            if (q >= end || (unsigned)(*q - '0') >= 10u) goto slow;
            while (q < end && (unsigned)(*q - '0') < 10u) {
                ex = ex * 10 + (*q - '0');
                if (ex > 10000) goto slow;
                ++q;
            }
            if (eneg) ex = -ex;
        }
        if (q != end) goto slow;
        ex -= frac;
        if (m >= ((uint64_t)1 << 53) || ex < -22 || ex > 22) goto slow;
        double v = (double)m;
        v = ex < 0 ? v / kPow10[-ex] : v * kPow10[ex];
        out = neg ? -v : v;
        return true;
    }
slow:
    double v = 0.0;
    auto res = std::from_chars(p, end, v);
    if (p == end || res.ec != std::errc() || res.ptr != end) return false;
    out = v;
    return true;
}

// One tree from text[b, e) (no terminating ';').  Same token semantics as Biopython's Newick parser:
// children in input order, '[...]' comments dropped, quoted labels, a ',' at top level creates a new root.
// The terminal counts are accumulated as clades are closed, and with `ladderize` the children of a clade are put in
// ladder order (ascending terminal count, stable: Bio.Phylo's ladderize, kamphir.py:154,280) the moment its ')' is read --
// the count of a clade does not depend on the order of anything, so sorting bottom-up while parsing gives the same
// tree as the reference's recursive top-down sort.
int parse_one(const char *text, size_t b, size_t e, RawTree &t, bool ladderize, std::vector<int32_t> &kids,
              std::string &err) {
    t.clear();
    t.reserve((e - b) / 12 + 16);
    int32_t cur = t.add(-1);
    t.root = cur;
    long lp = 0, rp = 0;
    bool synthetic_root = false;
    std::vector<RawNode> &nd = t.nd;
    auto close = [&](int32_t v) {            // v will get no more children: its terminal count is final
        if (nd[v].first < 0) nd[v].ntip = 1;
        if (nd[v].parent >= 0) nd[nd[v].parent].ntip += nd[v].ntip;
    };
    auto ladder = [&](int32_t v) {
        if (nd[v].nchild == 2) {             // binary node: swap only when strictly larger first (stable on ties)
            const int32_t x = nd[v].first, y = nd[x].next;
            if (nd[x].ntip > nd[y].ntip) {
                nd[v].first = y;
                nd[y].next = x;
                nd[x].next = -1;
                nd[v].last = x;
            }
        } else if (nd[v].nchild > 2) {
            kids.clear();
            for (int32_t c = nd[v].first; c >= 0; c = nd[c].next) kids.push_back(c);
            std::stable_sort(kids.begin(), kids.end(), [&](int32_t a, int32_t c) { return nd[a].ntip < nd[c].ntip; });
            nd[v].first = kids.front();
            nd[v].last = kids.back();
            for (size_t k = 0; k < kids.size(); ++k) nd[kids[k]].next = (k + 1 < kids.size()) ? kids[k + 1] : -1;
        }
    };
    size_t i = b;
    while (i < e) {
        char ch = text[i];
        if (is_space(ch)) { ++i; continue; }
        if (ch == '(') {
            cur = t.add(cur);
            ++lp; ++i;
        } else if (ch == ',') {
            int32_t par = nd[cur].parent;
            if (par < 0) {                 // outer parentheses missing: new root above the current clade
                int32_t nr = t.add(-1);
                t.link(nr, cur);
                t.root = nr;
                par = nr;
                synthetic_root = true;
            }
            close(cur);
            cur = t.add(par);
            ++i;
        } else if (ch == ')') {
            int32_t par = nd[cur].parent;
            if (par < 0) { err = "Newick: parenthesis mismatch"; return PK_ERR_PARSE; }
            close(cur);
            cur = par;
            if (ladderize) ladder(cur);
            ++rp; ++i;
        } else if (ch == ':') {
            size_t j = i + 1;
            while (j < e && text[j] == ' ') ++j;
            // one pass over the usual form digits[.digits][e[+-]digits] (see parse_length: exact when the mantissa has at
            // most 19 digits and the power of ten is within 10^22); anything else is delimited and handed to parse_length
            {
                static const double kP10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                                1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
                const char *q = text + j, *const end = text + e;
                uint64_t m = 0;
                int digits = 0, frac = 0, ex = 0;
                bool any = false, ok = true;
                while (q < end && (unsigned)(*q - '0') < 10u) {
                    if (m || *q != '0') { ++digits; m = m * 10 + (unsigned)(*q - '0'); }
                    any = true;
                    ++q;
                }
                if (q < end && *q == '.') {
                    ++q;
                    // eight fraction digits per step while they last (simulators write 10-12 of them): `digits` then
                    // counts 8 for a block that starts the mantissa with zeros -- an upper bound, which at worst sends
                    // a 12+-digit mantissa behind many zeros to parse_length (same value, slower)
                    while (end - q >= 8) {
                        uint64_t w;
                        std::memcpy(&w, q, 8);
                        if (!all_digits8(w)) break;
                        const uint32_t v8 = digits8(w);
                        if (m || v8) { digits += 8; if (digits <= 19) m = m * 100000000u + v8; }
                        any = true;
                        frac += 8;
                        q += 8;
                    }
                    while (q < end && (unsigned)(*q - '0') < 10u) {
                        if (m || *q != '0') { ++digits; if (digits <= 19) m = m * 10 + (unsigned)(*q - '0'); }
                        any = true;
                        ++frac;
                        ++q;
                    }
                }
                if (any && q < end && (*q == 'e' || *q == 'E')) {
                    const char *q0 = q++;
                    bool eneg = false;
                    if (q < end && (*q == '+' || *q == '-')) { eneg = *q == '-'; ++q; }
                    if (q < end && (unsigned)(*q - '0') < 10u) {
                        while (q < end && (unsigned)(*q - '0') < 10u && ex < 100000) ex = ex * 10 + (*q++ - '0');
                        if (eneg) ex = -ex;
                    } else {
                        q = q0;
                        ok = false;
                    }
                }
                ex -= frac;
                if (ok && any && digits <= 19 && m < ((uint64_t)1 << 53) && ex >= -22 && ex <= 22 &&
                    (q == end || is_token_end(*q))) {
                    const double v = (double)m;
                    nd[cur].bl = ex < 0 ? v / kP10[-ex] : v * kP10[ex];
                    i = (size_t)(q - text);
                    continue;
                }
            }
            size_t k = j;
            while (k < e && !is_token_end(text[k])) ++k;
            size_t s = j;
            if (s < k && text[s] == '+') ++s;
            double v = 0.0;
            if (s == k || !parse_length(text + s, text + k, v)) {
                err = "Newick: bad branch length '" + std::string(text + j, k - j) + "'";
                return PK_ERR_PARSE;
            }
            nd[cur].bl = v;
            i = k;
        } else if (ch == '\'') {
            size_t j = i + 1;
            while (j < e && text[j] != '\'') ++j;
            if (j >= e) { err = "Newick: unterminated quoted label"; return PK_ERR_PARSE; }
            nd[cur].name_off = (int32_t)(i + 1 - b);
            nd[cur].name_len = (int32_t)(j - i - 1);
            i = j + 1;
        } else if (ch == '[') {
            size_t j = i + 1;
            while (j < e && text[j] != ']') ++j;
            if (j >= e) { err = "Newick: unterminated comment"; return PK_ERR_PARSE; }
            i = j + 1;
        } else if (ch == ';') {
            break;
        } else if (ch == ']') {
            err = "Newick: unexpected ']'";
            return PK_ERR_PARSE;
        } else {
            size_t k = i;
            while (k < e && !is_token_end(text[k])) ++k;
            nd[cur].name_off = (int32_t)(i - b);
            nd[cur].name_len = (int32_t)(k - i);
            i = k;
        }
    }
    if (lp != rp) { err = "Newick: number of open/close parentheses do not match"; return PK_ERR_PARSE; }
    if (cur != t.root && !(synthetic_root && nd[cur].parent == t.root)) {
        err = "Newick: parenthesis mismatch";
        return PK_ERR_PARSE;
    }
    if (cur != t.root) {                   // synthetic root: its last child and the root itself were never closed by a ')'
        close(cur);
        if (ladderize) ladder(t.root);
    }
    if (nd[t.root].first < 0) nd[t.root].ntip = 1;       // a lone tip
    return PK_OK;
}

// splits a multi-tree text on ';' outside quotes / comments
void split_trees(const char *text, size_t len, std::vector<std::pair<size_t, size_t>> &out) {
    size_t start = 0, i = 0;
    auto flush = [&](size_t end) {
        size_t a = start, z = end;
        while (a < z && is_space(text[a])) ++a;
        while (z > a && is_space(text[z - 1])) --z;
        if (z > a) out.emplace_back(a, z);
    };
    if (!std::memchr(text, '\'', len) && !std::memchr(text, '[', len)) {   // no quotes, no comments: every ';' separates
        while (i < len) {
            const char *q = static_cast<const char *>(std::memchr(text + i, ';', len - i));
            if (!q) break;
            flush((size_t)(q - text));
            start = i = (size_t)(q - text) + 1;
        }
        flush(len);
        return;
    }
    while (i < len) {
        char ch = text[i];
        if (ch == '\'') {
            size_t j = i + 1;
            while (j < len && text[j] != '\'') ++j;
            i = (j < len) ? j + 1 : len;
            continue;
        }
        if (ch == '[') {
            size_t j = i + 1;
            while (j < len && text[j] != ']') ++j;
            i = (j < len) ? j + 1 : len;
            continue;
        }
        if (ch == ';') {
            flush(i);
            start = i + 1;
        }
        ++i;
    }
    flush(len);
}

int tip_state(const char *text, const RawTree &t, int32_t v, int nstates, int &state, std::string &err) {
    const char *p = text + t.nd[v].name_off;
    int len = t.nd[v].name_len;
    int k = len;
    while (k > 0 && p[k - 1] != '_') --k;          // token after the last '_'  (rcolgem.py:274)
    int val = 0;
    auto res = std::from_chars(p + k, p + len, val);
    // states are 0 .. S-1, or 1 .. S the way rcolgem numbers its demes (paste(1:n.tips, demes.sample, sep='_'),
    // rcolgem.py:274): S is taken as state 0, so either convention maps one-to-one onto [0, S)
    if (k == len || res.ec != std::errc() || res.ptr != p + len || val < 0 || val > nstates) {
        err = "tip label '" + std::string(p, len) + "' does not end in _<state> with 0 <= state <= " +
              std::to_string(nstates);
        return PK_ERR_INVALID;
    }
    state = val % nstates;
    return PK_OK;
}

// RawTree (already in ladder order if asked for) -> pk_tree following kamphir's prep order, in ONE depth-first
// traversal: on the way down the un-normalised depths and the pre-order sum of the branch lengths (the order Bio.Phylo's
// total_branch_length adds them in -- the mean must come out bit-identical), on the way up the post-order numbering of
// annotate_tree with the children's indices, classes and wavefront levels.
int flatten(const char *text, Scratch &sc, int flags, int normalize, int nstates, pk_tree &out, std::string &err) {
    RawTree &r = sc.raw;
    std::vector<RawNode> &nd = r.nd;
    const int32_t total = (int32_t)r.size();
    const int32_t root = r.root;
    if (normalize != PK_NORM_MEAN && normalize != PK_NORM_MEDIAN && normalize != PK_NORM_NONE) {
        err = "unknown normalize mode";
        return PK_ERR_INVALID;
    }
    if (flags & PK_PREP_ZERO_ROOT) nd[root].bl = 0.0;                      // kamphir.py:153,279

    const int32_t n_internal = total - nd[root].ntip;
    if (n_internal < 1) { err = "tree has no internal node (reference: IndexError at phyloK2.py:169)"; return PK_ERR_INVALID; }
    out.n = n_internal;
    out.ntips = nd[root].ntip;
    out.nstates = nstates;
    out.ncls = nstates > 0 ? 1 + 2 * nstates + nstates * nstates : 3;
    if (out.ncls > PK_MAX_CLASSES) { err = "too many tip states (max 7)"; return PK_ERR_INVALID; }
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    out.prod.assign(n_internal, 0);
    out.cls.assign(n_internal, 0);
    out.cidx.assign(2 * (size_t)n_internal, -1);
    out.bl.assign(2 * (size_t)n_internal, NaN);
    out.level.assign(n_internal, 0);
    out.heights.clear();
    out.heights.reserve((size_t)n_internal);

    std::vector<double> &depth = sc.depth;
    std::vector<int32_t> &pidx = sc.pidx;
    depth.resize(total);
    pidx.assign(total, -1);
    std::vector<double> &lens = sc.lens;
    lens.clear();
    double sum = 0.0, hmax = 0.0;
    bool firstv = true;
    int32_t next_idx = 0;
    const bool median = normalize == PK_NORM_MEDIAN;

    auto enter = [&](int32_t v) {                                          // pre-order visit
        const double bl = nd[v].bl;
        const double b = std::isnan(bl) ? 0.0 : bl;                        // depths: None counts as 0 (kamphir.py:118-119)
        depth[v] = (nd[v].parent >= 0 ? depth[nd[v].parent] : 0.0) + b;
        if (firstv || depth[v] > hmax) hmax = depth[v];
        firstv = false;
        if (!std::isnan(bl) && bl != 0.0) sum = sum + bl;                  // total_branch_length(): truthy lengths
        if (median && v != root) lens.push_back(bl);
        if (nd[v].first >= 0) out.heights.push_back(depth[v]);             // turned into heights below
    };
    auto leave = [&](int32_t v) -> int {                                   // post-order visit of an internal node
        const int32_t i = next_idx++;
        pidx[v] = i;
        if (nd[v].nchild != 2) {
            err = "internal node " + std::to_string(i) + " has " + std::to_string(nd[v].nchild) +
                  " children; the kernel is defined for strictly binary trees (reference: IndexError at phyloK2.py:185-189)";
            return PK_ERR_INVALID;
        }
        const int32_t kid[2] = {nd[v].first, nd[nd[v].first].next};
        int nterm = 0, st[2] = {-1, -1}, lv = 0;
        for (int s = 0; s < 2; ++s) {
            const int32_t c = kid[s];
            out.bl[2 * (size_t)i + s] = nd[c].bl;                          // rescaled and checked after the traversal
            if (nd[c].first < 0) {
                ++nterm;
                if (nstates > 0) {
                    int rc = tip_state(text, r, c, nstates, st[s], err);
                    if (rc != PK_OK) return rc;
                }
            } else {
                out.cidx[2 * (size_t)i + s] = pidx[c];
                lv = std::max(lv, out.level[pidx[c]] + 1);
            }
        }
        out.level[i] = lv;
        out.prod[i] = (uint8_t)(nterm + 1);                                 // phyloK2.py:141-142
        if (nstates == 0) {
            out.cls[i] = (uint8_t)nterm;
        } else if (nterm == 0) {
            out.cls[i] = 0;
        } else if (nterm == 1) {
            const int slot = st[0] >= 0 ? 0 : 1;
            out.cls[i] = (uint8_t)(1 + slot * nstates + st[slot]);
        } else {
            out.cls[i] = (uint8_t)(1 + 2 * nstates + st[0] * nstates + st[1]);
        }
        return PK_OK;
    };
    {
        std::vector<std::pair<int32_t, int32_t>> &st = sc.st;              // (node, next child to visit)
        st.clear();
        enter(root);
        st.emplace_back(root, nd[root].first);
        while (!st.empty()) {
            auto &top = st.back();
            if (top.second >= 0) {
                const int32_t c = top.second;
                top.second = nd[c].next;
                enter(c);
                if (nd[c].first >= 0) st.emplace_back(c, nd[c].first);
            } else {
                const int rc = leave(top.first);
                if (rc != PK_OK) return rc;
                st.pop_back();
            }
        }
    }
    // node heights for kcoal (kamphir.py:255-258), un-normalised; sorted on first use (pk_sorted_heights)
    out.height = hmax;
    for (double &h : out.heights) h = hmax - h;
    out.heights_sorted = false;

    // normalize_tree (phyloK2.py:101-130): every non-root branch divided by the mean / median
    out.scale = 1.0;
    if (normalize == PK_NORM_MEAN || median) {
        const int32_t nbranches = total - 1;                               // :109
        if (nbranches < 1) { err = "normalize_tree: tree has no branches"; return PK_ERR_INVALID; }
        double scale;
        if (!median) {
            const double rootbl = std::isnan(nd[root].bl) ? 0.0 : nd[root].bl;   // reference: TypeError on None (phyloK2.py:113)
            scale = (sum - rootbl) / nbranches;
        } else {
            for (double b : lens)
                if (std::isnan(b)) { err = "normalize_tree: branch without a length"; return PK_ERR_INVALID; }
            std::sort(lens.begin(), lens.end());
            const int32_t half = nbranches / 2;                            // py2 integer division, :123-127
            scale = (nbranches % 2 == 0) ? (lens[half - 1] + lens[half]) / 2. : lens[half];
        }
        if (!(scale != 0.0) || !std::isfinite(scale)) {
            err = "normalize_tree: zero or non-finite scale (reference: ZeroDivisionError)";
            return PK_ERR_INVALID;
        }
        for (double &b : out.bl) b /= scale;                               // every entry is a non-root branch; NaN stays NaN
        out.scale = scale;
    }
    for (double b : out.bl)
        if (!std::isfinite(b)) {
            err = std::isnan(b) ? "branch without a length (reference: TypeError at phyloK2.py:146)" : "non-finite branch length";
            return PK_ERR_INVALID;
        }
    return pk_tree_finish(out, /*levels_known=*/true);
}

}  // namespace

// levels, kernel order, per-class level table, child-class histograms
int pk_tree_finish(pk_tree &t, bool levels_known) {
    const int32_t n = t.n, ncls = t.ncls;
    int32_t nlev = 1;
    if (levels_known) {                    // the flattener computed them on its way up
        for (int32_t i = 0; i < n; ++i) nlev = std::max(nlev, t.level[i] + 1);
    } else {
        t.level.assign(n, 0);
        for (int32_t i = 0; i < n; ++i) {
            int32_t lv = 0;
            for (int s = 0; s < 2; ++s) {
                int32_t c = t.cidx[2 * (size_t)i + s];
                if (c >= 0) {
                    if (c >= i) return pk_fail(PK_ERR_INVALID, "child index is not below its parent in post-order");
                    lv = std::max(lv, t.level[c] + 1);
                }
            }
            t.level[i] = lv;
            nlev = std::max(nlev, lv + 1);
            if (t.cls[i] >= ncls) return pk_fail(PK_ERR_INVALID, "class id out of range");
        }
    }
    t.nlev = nlev;
    // temporaries of this function live per worker thread: a tree set is finished by the same few threads again and again
    struct FinishScratch {
        std::vector<int32_t> cnt, pos, parent, cur, nxt, fill;
        std::vector<uint32_t> keys;
    };
    static thread_local FinishScratch fs;
    // counting sort by (class, level); stable in post-order index
    std::vector<int32_t> &cnt = fs.cnt;
    cnt.assign((size_t)ncls * nlev + 1, 0);
    for (int32_t i = 0; i < n; ++i) cnt[(size_t)t.cls[i] * nlev + t.level[i] + 1]++;
    for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
    t.lvl.assign((size_t)ncls * (nlev + 1), 0);
    t.cls_start.assign(ncls + 1, 0);
    for (int32_t c = 0; c < ncls; ++c) {
        for (int32_t l = 0; l <= nlev; ++l) t.lvl[(size_t)c * (nlev + 1) + l] = cnt[(size_t)c * nlev + l];
        t.cls_start[c] = cnt[(size_t)c * nlev];
    }
    t.cls_start[ncls] = n;
    t.order.resize(n);
    t.rank.resize(n);
    std::vector<int32_t> &pos = fs.pos;
    pos.assign(cnt.begin(), cnt.end() - 1);
    for (int32_t i = 0; i < n; ++i) {
        int32_t row = pos[(size_t)t.cls[i] * nlev + t.level[i]]++;
        t.order[row] = i;
        t.rank[i] = row - t.cls_start[t.cls[i]];
    }
    // parent codes and the column order (class, parent code, ancestor path, post-order)
    const int npc = 2 * ncls + 2;
    t.pcode.assign(n, 0);
    std::vector<int32_t> &parent = fs.parent;
    parent.assign(n, -1);
    for (int32_t i = 0; i < n; ++i)
        for (int s = 0; s < 2; ++s) {
            int32_t c = t.cidx[2 * (size_t)i + s];
            if (c >= 0) {
                t.pcode[c] = (uint8_t)((t.cls[i] + 1) * 2 + s);
                parent[c] = i;
            }
        }
    t.gstart.assign((size_t)ncls * npc + 1, 0);
    for (int32_t i = 0; i < n; ++i) t.gstart[(size_t)t.cls[i] * npc + t.pcode[i] + 1]++;
    for (size_t k = 1; k < t.gstart.size(); ++k) t.gstart[k] += t.gstart[k - 1];
    // Within a (class, parent code) group the columns are ordered by the parent codes of their ancestors (parent first),
    // i.e. like their parents are ordered among the columns of the parents' class.  The child cells that 32 consecutive
    // columns gather from one stored row are then (almost) consecutive too: 3.6 sectors per gather instead of 6.8 at
    // 500 tips.  The own parent code plus those of four ancestors are enough (measured: 3.9 sectors with 4 keys, 3.6 with
    // 6 or 40).
    t.colorder.resize(n);
    t.krank.resize(n);
    {
        // The four ancestor keys of a node in one word, byte d = parent code of the ancestor d levels above the parent (0 beyond
        // the root; codes are below 2 * 64 + 2): a parent comes after its children in post-order, so one pass from the root down
        // builds every word from the parent's, keys(i) = pcode(parent) | keys(parent) << 8.
        std::vector<uint32_t> &keys = fs.keys;
        keys.resize(n);
        for (int32_t i = n - 1; i >= 0; --i) {
            const int32_t p = parent[i];
            if (p >= 0 && p <= i) return pk_fail(PK_ERR_INVALID, "child index is not below its parent in post-order");
            keys[i] = p >= 0 ? ((uint32_t)t.pcode[p] | (keys[p] << 8)) : 0u;
        }
        // least-significant-key-first radix sort, every pass a stable counting sort over at most npc buckets: parent code
        // of the 4th, 3rd, 2nd ancestor, of the parent, then the (class, parent code) group; ties keep the post-order
        std::vector<int32_t> &cur = fs.cur, &nxt = fs.nxt;
        cur.resize(n);
        nxt.resize(n);
        for (int32_t i = 0; i < n; ++i) cur[i] = i;
        for (int d = 3; d >= 0; --d) {
            const int sh = 8 * d;
            cnt.assign((size_t)npc + 1, 0);
            for (int32_t i = 0; i < n; ++i) cnt[((keys[i] >> sh) & 0xffu) + 1]++;
            if (cnt[1] == n) continue;                 // every node has key 0 at this depth (shallow tree): order unchanged
            for (int k = 0; k < npc; ++k) cnt[k + 1] += cnt[k];
            for (int32_t c = 0; c < n; ++c) nxt[cnt[(keys[cur[c]] >> sh) & 0xffu]++] = cur[c];
            cur.swap(nxt);
        }
        std::vector<int32_t> &fill = fs.fill;
        fill.assign(t.gstart.begin(), t.gstart.end() - 1);
        for (int32_t c = 0; c < n; ++c) {
            const int32_t i = cur[c];
            const size_t g = (size_t)t.cls[i] * npc + t.pcode[i];
            const int32_t col = fill[g]++;
            t.colorder[col] = i;
            t.krank[i] = col - t.gstart[g];
        }
    }
    t.hist.assign((size_t)ncls * 2 * (ncls + 1), 0);
    for (int32_t i = 0; i < n; ++i)
        for (int s = 0; s < 2; ++s) {
            int32_t c = t.cidx[2 * (size_t)i + s];
            int q = c < 0 ? 0 : t.cls[c] + 1;
            t.hist[((size_t)t.cls[i] * 2 + s) * (ncls + 1) + q]++;
        }
    return PK_OK;
}

// ---------------------------------------------------------------------------------------------------
// device block packing
// ---------------------------------------------------------------------------------------------------
static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

size_t pk_rowpart_bytes(const pk_tree &t, int /*precision*/) {
    size_t b = PK_HDR_INTS * 4;
    b += align16((size_t)(t.ncls + 1) * 4);
    b += align16((size_t)t.ncls * (t.nlev + 1) * 4);
    b += (size_t)t.n * 32;                                   // one 32-byte row record per internal node
    return b;
}

size_t pk_colpart_bytes(const pk_tree &t, int precision) {
    const size_t w = precision == 32 ? 4 : 8;
    const size_t npc = 2 * (size_t)t.ncls + 2;
    size_t b = PK_HDR_INTS * 4;
    b += align16((size_t)(t.ncls + 1) * 4);
    b += align16(((size_t)t.ncls * npc + 1) * 4);
    b += 2 * align16((size_t)t.n * w);
    b += 3 * align16((size_t)t.n * 4);
    return b;
}

static void put_real(unsigned char *dst, size_t w, int32_t k, double v) {
    if (w == 8) reinterpret_cast<double *>(dst)[k] = v;
    else reinterpret_cast<float *>(dst)[k] = (float)v;
}

void pk_rowpart_pack(const pk_tree &t, int precision, unsigned char *dst) {
    const size_t total = pk_rowpart_bytes(t, precision);
    std::memset(dst, 0, total);
    int32_t *hdr = reinterpret_cast<int32_t *>(dst);
    size_t off = PK_HDR_INTS * 4;
    hdr[0] = t.n; hdr[1] = t.ncls; hdr[2] = t.nlev; hdr[3] = (int32_t)total; hdr[11] = precision;
    for (int sl = 0; sl < 2; ++sl) {   // classes whose nodes all have a tip in slot sl (no child-cell gathers for that slot)
        uint64_t mask = 0;
        for (int c = 0; c < t.ncls; ++c) {
            bool all = true;
            for (int32_t row = t.cls_start[c]; row < t.cls_start[c + 1] && all; ++row)
                all = t.cidx[2 * (size_t)t.order[row] + sl] < 0;
            if (all) mask |= (uint64_t)1 << c;
        }
        hdr[7 + 2 * sl] = (int32_t)(mask & 0xffffffffu);
        hdr[8 + 2 * sl] = (int32_t)(mask >> 32);
    }
    hdr[4] = (int32_t)off;
    std::memcpy(dst + off, t.cls_start.data(), (size_t)(t.ncls + 1) * 4);
    off += align16((size_t)(t.ncls + 1) * 4);
    hdr[5] = (int32_t)off;
    std::memcpy(dst + off, t.lvl.data(), t.lvl.size() * 4);
    off += align16((size_t)t.ncls * (t.nlev + 1) * 4);
    hdr[6] = (int32_t)off;                                   // row records
    std::vector<int32_t> rowof(t.n);                         // post-order index -> kernel row
    for (int32_t row = 0; row < t.n; ++row) rowof[t.order[row]] = row;
    // record (32 bytes): a0, a1 (real), then 4 ints: kernel row of child 0, of child 1 (the kernel replaces them by the
    // children's C offsets), 0 (becomes the row's store offset), packed = q0 | q1 << 8 | parent code << 16 | class << 24
    // with q = class + 1 of the child, 0 for a tip
    for (int32_t row = 0; row < t.n; ++row) {
        const int32_t i = t.order[row];
        unsigned char *rec = dst + off + (size_t)row * 32;
        int32_t *ints;
        if (precision == 32) {
            reinterpret_cast<float *>(rec)[0] = (float)t.bl[2 * (size_t)i];
            reinterpret_cast<float *>(rec)[1] = (float)t.bl[2 * (size_t)i + 1];
            ints = reinterpret_cast<int32_t *>(rec + 16);
        } else {
            reinterpret_cast<double *>(rec)[0] = t.bl[2 * (size_t)i];
            reinterpret_cast<double *>(rec)[1] = t.bl[2 * (size_t)i + 1];
            ints = reinterpret_cast<int32_t *>(rec + 16);
        }
        int32_t packed = ((int32_t)t.pcode[i] << 16) | ((int32_t)t.cls[i] << 24);
        for (int s = 0; s < 2; ++s) {
            const int32_t c = t.cidx[2 * (size_t)i + s];
            ints[s] = c < 0 ? 0 : rowof[c];
            if (c >= 0) packed |= ((int32_t)t.cls[c] + 1) << (8 * s);
        }
        ints[2] = 0;
        ints[3] = packed;
    }
}

void pk_colpart_pack(const pk_tree &t, int precision, unsigned char *dst) {
    const size_t w = precision == 32 ? 4 : 8;
    const size_t total = pk_colpart_bytes(t, precision);
    const size_t npc = 2 * (size_t)t.ncls + 2;
    std::memset(dst, 0, total);
    int32_t *hdr = reinterpret_cast<int32_t *>(dst);
    size_t off = PK_HDR_INTS * 4;
    hdr[0] = t.n; hdr[1] = t.ncls; hdr[2] = t.nlev; hdr[3] = (int32_t)total; hdr[11] = precision;
    hdr[4] = (int32_t)off;
    std::memcpy(dst + off, t.cls_start.data(), (size_t)(t.ncls + 1) * 4);
    off += align16((size_t)(t.ncls + 1) * 4);
    hdr[5] = (int32_t)off;
    std::memcpy(dst + off, t.gstart.data(), t.gstart.size() * 4);
    off += align16(((size_t)t.ncls * npc + 1) * 4);
    for (int s = 0; s < 2; ++s) {
        hdr[6 + s] = (int32_t)off;
        for (int32_t col = 0; col < t.n; ++col) put_real(dst + off, w, col, t.bl[2 * (size_t)t.colorder[col] + s]);
        off += align16((size_t)t.n * w);
    }
    for (int s = 0; s < 2; ++s) {
        hdr[8 + s] = (int32_t)off;
        int32_t *code = reinterpret_cast<int32_t *>(dst + off);
        for (int32_t col = 0; col < t.n; ++col) {
            const int32_t c = t.cidx[2 * (size_t)t.colorder[col] + s];
            code[col] = c < 0 ? 0 : (((int32_t)t.cls[c] + 1) << 24) | t.krank[c];
        }
        off += align16((size_t)t.n * 4);
    }
    hdr[10] = (int32_t)off;
    int32_t *pc = reinterpret_cast<int32_t *>(dst + off);
    for (int32_t col = 0; col < t.n; ++col) pc[col] = t.pcode[t.colorder[col]];
}

// ---------------------------------------------------------------------------------------------------
// C ABI, host part
// ---------------------------------------------------------------------------------------------------
extern "C" int pk_tree_from_newick(const char *text, size_t len, int prep_flags, int normalize, int n_states,
                                   pk_tree **out) {
    if (!text || !out) return pk_fail(PK_ERR_INVALID, "null argument");
    if (n_states < 0) return pk_fail(PK_ERR_INVALID, "n_states < 0");
    *out = nullptr;
    std::vector<std::pair<size_t, size_t>> ranges;
    split_trees(text, len, ranges);
    if (ranges.size() != 1)
        return pk_fail(PK_ERR_PARSE, "expected exactly one tree, found " + std::to_string(ranges.size()));
    static thread_local Scratch sc;
    RawTree &raw = sc.raw;
    std::string err;
    int rc = parse_one(text, ranges[0].first, ranges[0].second, raw, (prep_flags & PK_PREP_LADDERIZE) != 0, sc.kids, err);
    if (rc != PK_OK) return pk_fail(rc, err);
    pk_tree *t = new (std::nothrow) pk_tree();
    if (!t) return pk_fail(PK_ERR_NOMEM, "out of memory");
    rc = flatten(text + ranges[0].first, sc, prep_flags, normalize, n_states, *t, err);
    if (rc != PK_OK) {
        delete t;
        return pk_fail(rc, err.empty() ? std::string(pk_last_error()) : err);
    }
    *out = t;
    return PK_OK;
}

// shared by the two bulk entry points: tree k is text_of(k)[ranges[k].first, ranges[k].second)
static int flatten_many(const std::vector<const char *> &bases, const std::vector<std::pair<size_t, size_t>> &ranges,
                        size_t total_len, int prep_flags, int normalize, int n_states, int nthreads, pk_tree ***out,
                        int32_t *n_out, const std::function<void(size_t, pk_tree &)> *post = nullptr);

extern "C" int pk_trees_from_newick(const char *text, size_t len, int prep_flags, int normalize, int n_states,
                                    int nthreads, pk_tree ***out, int32_t *n_out) {
    if (!text || !out || !n_out) return pk_fail(PK_ERR_INVALID, "null argument");
    *out = nullptr;
    *n_out = 0;
    std::vector<std::pair<size_t, size_t>> ranges;
    split_trees(text, len, ranges);
    std::vector<const char *> bases(ranges.size(), text);
    return flatten_many(bases, ranges, len, prep_flags, normalize, n_states, nthreads, out, n_out);
}

extern "C" int pk_trees_from_newick_list(const char *const *texts, const size_t *lens, int32_t n, int prep_flags,
                                         int normalize, int n_states, int nthreads, pk_tree ***out, int32_t *n_out) {
    return pk_trees_from_newick_list_post(texts, lens, n, prep_flags, normalize, n_states, nthreads, out, n_out, nullptr);
}

int pk_trees_from_newick_list_post(const char *const *texts, const size_t *lens, int32_t n, int prep_flags, int normalize,
                                   int n_states, int nthreads, pk_tree ***out, int32_t *n_out,
                                   const std::function<void(size_t, pk_tree &)> *post) {
    if (!texts || !lens || !out || !n_out) return pk_fail(PK_ERR_INVALID, "null argument");
    *out = nullptr;
    *n_out = 0;
    std::vector<const char *> bases((size_t)std::max(n, 0));
    std::vector<std::pair<size_t, size_t>> ranges((size_t)std::max(n, 0));
    size_t total = 0;
    for (int32_t k = 0; k < n; ++k) {
        if (!texts[k]) return pk_fail(PK_ERR_INVALID, "null Newick string in list");
        bases[k] = texts[k];
        size_t e = lens[k];
        while (e > 0 && (is_space(texts[k][e - 1]) || texts[k][e - 1] == ';')) --e;   // one tree per string
        ranges[k] = std::make_pair((size_t)0, e);
        total += lens[k];
    }
    return flatten_many(bases, ranges, total, prep_flags, normalize, n_states, nthreads, out, n_out, post);
}

static int flatten_many(const std::vector<const char *> &bases, const std::vector<std::pair<size_t, size_t>> &ranges,
                        size_t len, int prep_flags, int normalize, int n_states, int nthreads, pk_tree ***out,
                        int32_t *n_out, const std::function<void(size_t, pk_tree &)> *post) {
    const size_t nt = ranges.size();
    pk_tree **arr = new (std::nothrow) pk_tree *[nt + 1]();
    if (!arr) return pk_fail(PK_ERR_NOMEM, "out of memory");
    if (nthreads <= 0) nthreads = pk_host_threads();
    // a worker per ~16 KB of text (waking a parked pool thread costs about as much as parsing that)
    nthreads = (int)std::min<size_t>((size_t)nthreads, std::max<size_t>(1, std::min<size_t>(nt, len / (16 * 1024))));
    std::atomic<size_t> next{0};
    std::atomic<int> status{PK_OK};
    std::string first_err;
    std::atomic<bool> have_err{false};
    auto worker = [&](int) {
        static thread_local Scratch sc;
        RawTree &raw = sc.raw;
        std::string err;
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= nt || status.load() != PK_OK) break;
            const char *text = bases[k];
            int rc = parse_one(text, ranges[k].first, ranges[k].second, raw, (prep_flags & PK_PREP_LADDERIZE) != 0, sc.kids, err);
            if (rc == PK_OK) {
                arr[k] = new (std::nothrow) pk_tree();
                if (!arr[k]) { rc = PK_ERR_NOMEM; err = "out of memory"; }
                else rc = flatten(text + ranges[k].first, sc, prep_flags, normalize, n_states, *arr[k], err);
                if (rc == PK_OK && post) (*post)(k, *arr[k]);      // per-tree follow-up work on the worker that has the tree in cache
            }
            if (rc != PK_OK) {
                bool expected = false;
                if (have_err.compare_exchange_strong(expected, true)) {
                    first_err = "tree " + std::to_string(k) + ": " + (err.empty() ? std::string(pk_last_error()) : err);
                    status.store(rc);
                }
                break;
            }
        }
    };
    pk_parallel(nthreads, worker);
    if (status.load() != PK_OK) {
        for (size_t k = 0; k < nt; ++k) delete arr[k];
        delete[] arr;
        return pk_fail(status.load(), first_err);
    }
    *out = arr;
    *n_out = (int32_t)nt;
    return PK_OK;
}

extern "C" void pk_tree_array_free(pk_tree **arr) { delete[] arr; }

extern "C" int pk_tree_from_arrays(int32_t n_internal, const int32_t *cidx, const double *bl, const uint8_t *cls,
                                   int32_t n_classes, pk_tree **out) {
    if (!cidx || !bl || !out) return pk_fail(PK_ERR_INVALID, "null argument");
    if (n_internal < 1) return pk_fail(PK_ERR_INVALID, "tree has no internal node (reference: IndexError at phyloK2.py:169)");
    *out = nullptr;
    pk_tree *t = new (std::nothrow) pk_tree();
    if (!t) return pk_fail(PK_ERR_NOMEM, "out of memory");
    t->n = n_internal;
    t->ncls = cls ? n_classes : 3;
    if (t->ncls < 1 || t->ncls > PK_MAX_CLASSES) {
        delete t;
        return pk_fail(PK_ERR_INVALID, "n_classes out of range");
    }
    t->nstates = cls ? -1 : 0;
    t->cidx.assign(cidx, cidx + 2 * (size_t)n_internal);
    t->bl.assign(bl, bl + 2 * (size_t)n_internal);
    t->prod.assign(n_internal, 0);
    t->cls.assign(n_internal, 0);
    int32_t tips = 0;
    for (int32_t i = 0; i < n_internal; ++i) {
        int nterm = 0;
        for (int s = 0; s < 2; ++s) {
            int32_t c = cidx[2 * (size_t)i + s];
            if (c < -1 || c >= n_internal) {
                delete t;
                return pk_fail(PK_ERR_INVALID, "child index out of range");
            }
            if (!std::isfinite(bl[2 * (size_t)i + s])) {
                delete t;
                return pk_fail(PK_ERR_INVALID, "missing or non-finite branch length");
            }
            if (c < 0) ++nterm;
        }
        tips += nterm;
        t->prod[i] = (uint8_t)(nterm + 1);
        t->cls[i] = cls ? cls[i] : (uint8_t)nterm;
    }
    t->ntips = tips;
    int rc = pk_tree_finish(*t);
    if (rc != PK_OK) {
        delete t;
        return rc;
    }
    *out = t;
    return PK_OK;
}

extern "C" void pk_tree_free(pk_tree *t) { delete t; }

// The facts a tree built from arrays lacks (it has no Newick text to take them from): node heights for kcoal
// (kamphir.py:255-258), the un-normalised height, the normalisation divisor and the tip count.  Used by the Python mirror
// to rebuild a pickled FlatTree completely (Kamphir pickles the whole object into Pool workers, kamphir.py:27-32).
extern "C" int pk_tree_set_heights(pk_tree *t, const double *heights, int32_t n, double height, double scale) {
    if (!t || !heights) return pk_fail(PK_ERR_INVALID, "null argument");
    if (n != t->n) return pk_fail(PK_ERR_INVALID, "one height per internal node expected");
    for (int32_t i = 0; i < n; ++i)
        if (!std::isfinite(heights[i])) return pk_fail(PK_ERR_INVALID, "non-finite node height");
    t->heights.assign(heights, heights + n);
    t->heights_sorted = false;
    t->height = height;
    t->scale = scale;
    return PK_OK;
}

extern "C" int pk_tree_get_info(const pk_tree *t, pk_tree_info *info) {
    if (!t || !info) return pk_fail(PK_ERR_INVALID, "null argument");
    info->n_internal = t->n;
    info->n_tips = t->ntips;
    info->n_levels = t->nlev;
    info->n_classes = t->ncls;
    info->n_states = t->nstates;
    info->reserved = 0;
    info->height = t->height;
    info->scale = t->scale;
    return PK_OK;
}

extern "C" int pk_tree_export(const pk_tree *t, uint8_t *prod, uint8_t *cprod, int32_t *cidx, double *bl) {
    if (!t) return pk_fail(PK_ERR_INVALID, "null argument");
    for (int32_t i = 0; i < t->n; ++i) {
        if (prod) prod[i] = t->nstates != 0 ? (uint8_t)(t->cls[i] + 1) : t->prod[i];
        for (int s = 0; s < 2; ++s) {
            int32_t c = t->cidx[2 * (size_t)i + s];
            if (cprod) cprod[2 * (size_t)i + s] = c < 0 ? 0 : (t->nstates != 0 ? (uint8_t)(t->cls[c] + 1) : t->prod[c]);
            if (cidx) cidx[2 * (size_t)i + s] = c;
            if (bl) bl[2 * (size_t)i + s] = t->bl[2 * (size_t)i + s];
        }
    }
    return PK_OK;
}

// ascending node heights (kamphir.py:257-258 sorts them); sorted on first use, under the tree's once-flag
static const std::vector<double> &pk_sorted_heights(const pk_tree *t) {
    pk_tree *m = const_cast<pk_tree *>(t);
    std::call_once(m->heights_once, [m]() {
        if (!m->heights_sorted) std::sort(m->heights.begin(), m->heights.end());
        m->heights_sorted = true;
    });
    return t->heights;
}

void pk_tree_sort_heights(const pk_tree *t) { pk_sorted_heights(t); }

extern "C" int pk_tree_node_heights(const pk_tree *t, double *heights) {
    if (!t || !heights) return pk_fail(PK_ERR_INVALID, "null argument");
    if ((int32_t)t->heights.size() != t->n) return pk_fail(PK_ERR_INVALID, "tree was built from arrays: no node heights");
    std::memcpy(heights, pk_sorted_heights(t).data(), sizeof(double) * (size_t)t->n);
    return PK_OK;
}

extern "C" int pk_pair_counts(const pk_tree *a, const pk_tree *b, int64_t counts[3]) {
    if (!a || !b || !counts) return pk_fail(PK_ERR_INVALID, "null argument");
    if (a->ncls != b->ncls) return pk_fail(PK_ERR_INVALID, "trees use different production class schemes");
    int64_t m = 0, r = 0;
    const int ncls = a->ncls;
    for (int c = 0; c < ncls; ++c) {
        m += (int64_t)(a->cls_start[c + 1] - a->cls_start[c]) * (b->cls_start[c + 1] - b->cls_start[c]);
        for (int s = 0; s < 2; ++s)
            for (int q = 1; q <= ncls; ++q)
                r += a->hist[((size_t)c * 2 + s) * (ncls + 1) + q] * b->hist[((size_t)c * 2 + s) * (ncls + 1) + q];
    }
    counts[0] = (int64_t)a->n * b->n;
    counts[1] = m;
    counts[2] = r;
    return PK_OK;
}

extern "C" int pk_kcoal(const pk_tree *sim, const pk_tree *target, double *out) {
    if (!sim || !target || !out) return pk_fail(PK_ERR_INVALID, "null argument");
    if ((int32_t)sim->heights.size() != sim->n || (int32_t)target->heights.size() != target->n)
        return pk_fail(PK_ERR_INVALID, "node heights unavailable (tree built from arrays)");
    if (sim->n > target->n)
        return pk_fail(PK_ERR_INVALID, "simulated tree has more internal nodes than the target (reference: IndexError at kamphir.py:264)");
    const std::vector<double> &hs = pk_sorted_heights(sim), &ht = pk_sorted_heights(target);
    double dot = 0, n1 = 0, n2 = 0;                                          // kamphir.py:260-269
    for (int32_t k = 0; k < sim->n; ++k) {
        double h = hs[k], g = ht[k];
        dot += h * g;
        n1 += h * h;
        n2 += g * g;
    }
    // all heights zero on one side: the reference divides by zero here (ZeroDivisionError at kamphir.py:269)
    if (!(n1 > 0.0) || !(n2 > 0.0))
        return pk_fail(PK_ERR_INVALID, "kcoal undefined: all node heights are zero (reference: ZeroDivisionError at kamphir.py:269)");
    *out = dot / (std::sqrt(n1) * std::sqrt(n2));
    return PK_OK;
}

extern "C" int pk_kcoal_batch(const pk_tree *const *sims, int32_t nsims, const pk_tree *target, double *out) {
    if (!sims || !target || (nsims > 0 && !out)) return pk_fail(PK_ERR_INVALID, "null argument");
    for (int32_t s = 0; s < nsims; ++s)
        if (!sims[s]) return pk_fail(PK_ERR_INVALID, "null tree in set");
    // the replicates' node heights are sorted here, on first use (kamphir.py:257-258): one worker per ~4 k nodes
    size_t nodes = 0;
    for (int32_t s = 0; s < nsims; ++s) nodes += (size_t)sims[s]->n;
    const int workers = (int)std::min<size_t>((size_t)pk_host_threads(), std::max<size_t>(1, nodes / 4096));
    std::atomic<int32_t> next{0};
    std::atomic<int> status{PK_OK};
    std::string first_err;
    std::mutex err_lock;
    pk_sorted_heights(target);
    pk_parallel(workers, [&](int) {
        for (;;) {
            const int32_t s = next.fetch_add(1);
            if (s >= nsims) break;
            int rc = pk_kcoal(sims[s], target, out + s);
            if (rc) {
                std::lock_guard<std::mutex> lk(err_lock);
                if (status.load() == PK_OK) {
                    first_err = std::string(pk_last_error());
                    status.store(rc);
                }
            }
        }
    });
    if (status.load() != PK_OK) return pk_fail(status.load(), "replicate: " + first_err);
    return PK_OK;
}
