// internal.cuh -- objects and helpers shared by the translation units of libhegpu.so
// (hegpu.cu: C ABI + composites; ntt_inst_*.cu: the NTT kernel instantiations of one job each,
// compiled in parallel).
#pragma once
#include "../../include/hegpu.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.cuh"

using namespace hegpu;

// ------------------------------------------------------------------------- errors
int fail(int code, const std::string &msg);  // records the calling thread's message (hegpu.cu)
#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            if (_e == cudaErrorMemoryAllocation) return fail(HEGPU_ERR_OUT_OF_MEMORY, std::string("out of device memory: ") + #expr); \
            return fail(HEGPU_ERR_CUDA, std::string(cudaGetErrorString(_e)) + " at " + #expr);    \
        }                                                                                         \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int _s = (expr);          \
        if (_s != HEGPU_OK) return _s; \
    } while (0)
#define INVALID(msg) return fail(HEGPU_ERR_INVALID_ARGUMENT, msg)
#define LOGIC(msg) return fail(HEGPU_ERR_LOGIC, msg)


// ------------------------------------------------------------------------- objects
struct Arena {  // grow-only device scratch, bump-allocated per API call
    char *base = nullptr;
    size_t cap = 0, off = 0;
};

enum ProfKind {
    PK_NTT_FWD_PLAIN, PK_NTT_INV_PLAIN, PK_KS_INTT, PK_KS_LIFT_NTT, PK_KS_INNER, PK_HALF_INTT, PK_KS_MODDOWN_NTT,
    PK_RESCALE_NTT, PK_BSGS_INNER, PK_TENSOR, PK_ELEMENTWISE, PK_DH_INNER, PK_COUNT
};
static const char *const kProfNames[PK_COUNT] __attribute__((unused)) = { "ntt_fwd_plain", "ntt_inv_plain", "ks_intt", "ks_lift_ntt", "ks_inner",
                                                  "half_intt", "ks_moddown_ntt", "rescale_ntt", "bsgs_inner", "tensor",
                                                  "elementwise", "dh_inner" };
struct ProfRec {
    int kind;
    cudaEvent_t a, b;
    u64 units, bytes;
};

struct hegpu_ctx {
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[PK_COUNT] = {};
    u64 prof_launches[PK_COUNT] = {}, prof_units[PK_COUNT] = {}, prof_bytes[PK_COUNT] = {};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_h2d = nullptr, copy_d2h = nullptr;  // async host<->device copies overlap compute
    cudaEvent_t ev_fence = nullptr;
    u32 n = 0, logn = 0, K = 0;
    std::vector<u64> q, psi;
    std::vector<int> level_bits;  // total_coeff_modulus_bit_count for L = 1..K
    // device tables
    ulonglong2 *d_fwd = nullptr, *d_inv = nullptr, *d_inv_last = nullptr;
    double *d_fwd_d = nullptr, *d_inv_d = nullptr;
    ModF64 *d_modsd = nullptr;
    ModConst *d_mods = nullptr;
    MdConst *d_md = nullptr;  // [K][K]: row d = dropped modulus, column i = target limb
    NttTables tabs{};
    // keys
    u64 *relin_key = nullptr;
    std::map<u32, u64 *> galois_keys;
    std::map<u32, u32 *> perms;
    Arena arena;
    u64 *park = nullptr;  // [jobs][N/2] scratch of the N = 16384 park kernels (L2-resident in practice)
    size_t park_words = 0;
    // second execution slot: composites run alternate batch chunks on `stream` and `aux_stream` so that
    // small grids of one chunk overlap the kernels of the other; each slot has its own park scratch
    static constexpr int MAX_SLOTS = 4;
    cudaStream_t main_stream = nullptr, aux_stream[MAX_SLOTS] = {};  // aux_stream[0] unused (slot 0 = main)
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_SLOTS] = {};
    u64 *park_slots[MAX_SLOTS] = {};
    size_t park_words_slots[MAX_SLOTS] = {};
    int cur_slot = 0;
    int n_slots = 2;  // HEGPU_STREAMS=k (1..4); 1 disables
    int use_park = 1;
    u64 *stage = nullptr;  // host<->device staging
    size_t stage_words = 0;
    u64 launches = 0;
    int sms = 148;
    int dh_f64 = 1;    // fused kernel: FP64-pipe arithmetic on the limbs whose modulus is below 2^40 (HEGPU_DH_F64=0: integer everywhere)
    int dh_imma = 0;   // HEGPU_DH_IMMA=1: every double-hoisted matvec takes the integer-MMA inner sums (same as the HEGPU_MATVEC_IMMA flag)
    int imma_tx = 8;   // coefficients per CTA of the integer-MMA kernel: 8 = two CTAs of 8 warps per SM (measured faster), HEGPU_IMMA_TX=16 = one CTA of 16 warps
    u64 key_epoch = 0; // bumped whenever a Galois key is (re)loaded: invalidates pre-multiplied diagonals
    int limb_major = 0;  // HEGPU_LIMB_MAJOR=1: mod-down / final NTT jobs in limb-major order (co-resident CTAs share one arithmetic policy)
    int dh_bk = 1;     // fused kernel, more than 4 giant steps: keep the rotated ciphertexts b_k of the first launch in HBM for the later ones (HEGPU_DH_BK=0: recompute)
    int dh_stcs = 1;   // fused kernel: evict-first stores of the inner sums (HEGPU_DH_STCS=0: plain stores)
    int dh_swz = 0;    // fused kernel: limb order rotated per wave of CTAs so that co-resident CTAs mix the two policies (HEGPU_DH_SWZ=1)
    int fuse_final = 1;  // double-hoisted matvec: final mod-down and rescale as one pass (HEGPU_FUSE_FINAL=0: two steps)
    int park32k = 2;     // N = 32768: 2 = two-level park (64 KiB CTAs, three per SM, the N = 16384 sub-transforms); HEGPU_PARK32K=1: one-level
                         // park in 128 KiB (one CTA per SM); 0: two CTAs per transform + finishing pass
    int dh_fused = 1;  // double-hoisted matvec: fused baby-step + inner-sum kernel (HEGPU_DH_FUSED=0: unfused kernels)
    int loge = 3;  // NTT register-set size at N = 16384: 3 = radix-8 passes, 256 threads x 80 registers, 3 CTAs per SM (HEGPU_LOGE=4: radix-16, 2 CTAs)
    size_t ws_budget = (size_t)24 << 30;  // scratch budget per composite chunk
    // kernels whose dynamic shared-memory limit this context has raised (kernel address -> bytes); per context, so
    // that contexts driven from different host threads never share mutable state
    std::map<const void *, size_t> smem_configured;
    // temporaries of the loop-order-exact composites, cached by (slot, batch, size_cap) and grown on demand: the
    // composites stay asynchronous on the context stream (no cudaMalloc / cudaFree / synchronize per call)
    std::map<std::tuple<int, u32, u32>, struct hegpu_ct *> tmp_cache;
};
int configure_smem(hegpu_ctx *c, const void *kernel, size_t bytes);  // raise a kernel's dynamic smem limit once

struct hegpu_ct {
    hegpu_ctx *ctx;
    u64 *d;
    u32 batch, size_cap, L_cap;
    u32 size, L;
    double scale;
    cudaEvent_t ev_copy = nullptr;      // last asynchronous host<->device copy of this batch
    mutable bool copy_pending = false;  // compute that touches the batch must wait for ev_copy first
    CtView view() const { return CtView{ d, (size_t)size_cap * L_cap * ctx->n, (size_t)L_cap * ctx->n, (size_t)ctx->n }; }
    CtView view_at(u32 b0) const
    {
        CtView v = view();
        v.p += b0 * v.sb;
        return v;
    }
};

struct hegpu_pt {
    hegpu_ctx *ctx;
    u64 *d;
    u32 count, L_cap, L;
    double scale;
    u64 *d_mont = nullptr;  // lazily built copy in Montgomery form (matvec diagonals)
    bool mont_valid = false;
    bool ext = false;       // limb L holds the residues mod the special prime (hegpu_pt_upload_ext)
    // HEGPU_MATVEC_IMMA (opt-in): diagonals pre-multiplied with the baby-step keys, as 8-bit-limb MMA fragments; valid for
    // one (n1, n2, g_first, L, key set) at a time
    u32 *d_w = nullptr;
    size_t w_words = 0;
    std::vector<u64> w_key;
    size_t stride() const { return (size_t)L_cap * ctx->n; }
};

// brackets one launch with events while profiling is enabled
struct Prof {
    hegpu_ctx *c;
    ProfRec r{};
    bool on;
    Prof(hegpu_ctx *c_, int kind, u64 units, u64 bytes) : c(c_), on(c_->profiling)
    {
        if (!on) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        r.kind = kind;
        r.units = units;
        r.bytes = bytes;
        r.a = get();
        r.b = get();
        cudaEventRecord(r.a, c->stream);
    }
    ~Prof()
    {
        if (!on) return;
        cudaEventRecord(r.b, c->stream);
        c->prof.push_back(r);
    }
};


int set_device(hegpu_ctx *c);
int arena_reserve(hegpu_ctx *c, size_t bytes);
int park_reserve(hegpu_ctx *c, size_t words);
int stage_reserve(hegpu_ctx *c, size_t words);
int ew_grid(hegpu_ctx *c, size_t total);
void select_slot(hegpu_ctx *c, int slot);  // make `slot` (0 = main stream, k = aux stream k) the one launches go to

struct ArenaPlan {  // first pass sizes the scratch, second pass hands out pointers
    hegpu_ctx *c;
    size_t off = 0;
    u64 *take(size_t words)
    {
        size_t bytes = (words * sizeof(u64) + 255) & ~(size_t)255;
        if (off + bytes > c->arena.cap) {  // a sizing bug: fail loudly instead of writing past the arena
            fprintf(stderr, "hegpu: scratch arena overrun (%zu + %zu > %zu bytes)\n", off, bytes, c->arena.cap);
            abort();
        }
        u64 *p = (u64 *)(c->arena.base + off);
        off += bytes;
        return p;
    }
};
static inline size_t align256(size_t words) { return ((words * sizeof(u64) + 255) & ~(size_t)255); }
