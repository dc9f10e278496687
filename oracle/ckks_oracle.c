/*
 * ckks_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see ckks_oracle.h).
 * PARITY UNPINNED against real SEAL 4.1 (not installable here); pinned only
 * against the known-answer prime chains of SURVEY.md 9.1, algebraic
 * identities and an independent Python big-integer restatement (tests/).
 *
 * Every function names the SEAL 4.1 routine / SURVEY.md section it restates and
 * the reference call site that reaches it.
 */
#include "ckks_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

#define ORC_MAXK 64

struct orc_ctx {
    u32 n, logn, K;
    u64 q[ORC_MAXK];
    u64 psi[ORC_MAXK];
    u64 ninv[ORC_MAXK], ninv_sh[ORC_MAXK];
    u64 r0[ORC_MAXK], r1[ORC_MAXK]; /* floor(2^128 / q) low / high word (SEAL Modulus::const_ratio) */
    u64 *w[ORC_MAXK];    /* psi^{brev(i)}            i in [0,N) */
    u64 *wsh[ORC_MAXK];  /* Shoup quotients of w     */
    u64 *iw[ORC_MAXK];   /* (psi^{brev(i)})^{-1}     */
    u64 *iwsh[ORC_MAXK];
};

/* ---------------------------------------------------------------- modular */
static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static inline u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }
static u64 powmod(u64 a, u64 e, u64 q)
{
    u64 r = 1;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
static inline u64 invmod(u64 a, u64 q) { return powmod(a, q - 2, q); }
static inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
/* x * w mod q in [0, 2q) */
static inline u64 mul_shoup_lazy(u64 x, u64 w, u64 wsh, u64 q)
{
    u64 hi = (u64)(((u128)x * wsh) >> 64);
    return x * w - hi * q;
}
static inline u32 brev(u32 x, u32 bits)
{
    u32 r = 0;
    for (u32 i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

/* SEAL util::barrett_reduce_128 / barrett_reduce_64 with the modulus' const_ratio */
static inline u64 barrett128(u128 z, u64 q, u64 r0, u64 r1)
{
    const u64 lo = (u64)z, hi = (u64)(z >> 64);
    u64 carry = (u64)(((u128)lo * r0) >> 64);
    u128 t2 = (u128)lo * r1;
    u64 tmp1 = (u64)t2 + carry;
    u64 tmp3 = (u64)(t2 >> 64) + (tmp1 < carry);
    t2 = (u128)hi * r0;
    u64 s = tmp1 + (u64)t2;
    carry = (u64)(t2 >> 64) + (s < tmp1);
    u64 qhat = hi * r1 + tmp3 + carry;
    u64 r = lo - qhat * q;
    r -= (r >= q) ? q : 0;
    r -= (r >= q) ? q : 0;
    return r;
}
static inline u64 barrett64(u64 x, u64 q, u64 r1)
{
    u64 h = (u64)(((u128)x * r1) >> 64);
    u64 r = x - h * q;
    return r >= q ? r - q : r;
}
#define MULMOD(c, i, a, b) barrett128((u128)(a) * (b), (c)->q[i], (c)->r0[i], (c)->r1[i])
#define REDUCE64(c, i, x) barrett64((x), (c)->q[i], (c)->r1[i])

/* ---------------------------------------------------------------- primes */
/* deterministic Miller-Rabin for 64-bit (SEAL util::is_prime is probabilistic; same set) */
int orc_is_prime(u64 n)
{
    if (n < 2) return 0;
    static const u64 small[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
    for (unsigned i = 0; i < 12; ++i) {
        if (n == small[i]) return 1;
        if (n % small[i] == 0) return 0;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) { d >>= 1; ++r; }
    for (unsigned i = 0; i < 12; ++i) {
        u64 x = powmod(small[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int j = 1; j < r; ++j) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

/* SEAL util::get_primes(factor, bit_size, count) -- numth.cpp; SURVEY 9.1 */
int orc_get_primes(u64 factor, int bits, u32 count, u64 *out)
{
    u64 value = ((((u64)1) << bits) - 1) / factor * factor + 1;
    u64 lower = ((u64)1) << (bits - 1);
    u32 got = 0;
    while (got < count && value > lower) {
        if (orc_is_prime(value)) out[got++] = value;
        value -= factor;
    }
    return got == count ? 0 : -1;
}

/* SEAL CoeffModulus::Create -- modulus.cpp; SURVEY 9.1.  Primes of one bit size are
 * generated descending and handed out from the back (smallest first). */
int orc_coeff_modulus_create(u32 n, const int *bits, u32 count, u64 *out)
{
    u32 used[65];
    u32 total[65];
    memset(used, 0, sizeof used);
    memset(total, 0, sizeof total);
    for (u32 i = 0; i < count; ++i) {
        if (bits[i] < 2 || bits[i] > 60) return -1;
        total[bits[i]]++;
    }
    u64 tmp[ORC_MAXK];
    for (u32 i = 0; i < count; ++i) {
        int b = bits[i];
        if (orc_get_primes(2ull * n, b, total[b], tmp)) return -1;
        out[i] = tmp[total[b] - 1 - used[b]];
        used[b]++;
    }
    return 0;
}

/* ---------------------------------------------------------------- context */
/* SEAL util::try_minimal_primitive_root(2N, q): the numerically smallest
 * primitive 2N-th root of unity -- numth.cpp; SURVEY 9.2 */
static u64 minimal_primitive_root(u64 q, u32 n)
{
    u64 two_n = 2ull * n;
    u64 e = (q - 1) / two_n;
    u64 root = 0;
    for (u64 g = 2;; ++g) {
        u64 r = powmod(g, e, q);
        if (powmod(r, n, q) == q - 1) { root = r; break; }
    }
    u64 sq = mulmod(root, root, q);
    u64 cur = root, best = root;
    for (u32 i = 0; i < n; ++i) { /* all odd powers = all primitive 2N-th roots */
        if (cur < best) best = cur;
        cur = mulmod(cur, sq, q);
    }
    return best;
}

orc_ctx *orc_ctx_create(u32 n, const u64 *moduli, u32 K)
{
    if (K == 0 || K > ORC_MAXK || n < 2 || (n & (n - 1))) return NULL; /* n = 2 admits the known answers of SEAL's own NTT unit test */
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->n = n;
    c->K = K;
    c->logn = 0;
    while ((1u << c->logn) < n) c->logn++;
    for (u32 i = 0; i < K; ++i) {
        u64 q = moduli[i];
        if (!orc_is_prime(q) || (q - 1) % (2ull * n)) { orc_ctx_free(c); return NULL; }
        c->q[i] = q;
        c->psi[i] = minimal_primitive_root(q, n);
        u64 ipsi = invmod(c->psi[i], q);
        c->w[i] = (u64 *)malloc(sizeof(u64) * n);
        c->wsh[i] = (u64 *)malloc(sizeof(u64) * n);
        c->iw[i] = (u64 *)malloc(sizeof(u64) * n);
        c->iwsh[i] = (u64 *)malloc(sizeof(u64) * n);
        u64 p = 1, ip = 1;
        for (u32 k = 0; k < n; ++k) { /* power k stored at brev(k) */
            u32 r = brev(k, c->logn);
            c->w[i][r] = p;
            c->wsh[i][r] = shoup(p, q);
            c->iw[i][r] = ip;
            c->iwsh[i][r] = shoup(ip, q);
            p = mulmod(p, c->psi[i], q);
            ip = mulmod(ip, ipsi, q);
        }
        {
            u128 ratio = (~(u128)0) / q; /* == floor(2^128 / q) for odd q */
            c->r0[i] = (u64)ratio;
            c->r1[i] = (u64)(ratio >> 64);
        }
        c->ninv[i] = invmod(n % q, q);
        c->ninv_sh[i] = shoup(c->ninv[i], q);
    }
    return c;
}

void orc_ctx_free(orc_ctx *c)
{
    if (!c) return;
    for (u32 i = 0; i < ORC_MAXK; ++i) {
        free(c->w[i]);
        free(c->wsh[i]);
        free(c->iw[i]);
        free(c->iwsh[i]);
    }
    free(c);
}
u32 orc_n(const orc_ctx *c) { return c->n; }
u32 orc_K(const orc_ctx *c) { return c->K; }
u64 orc_modulus(const orc_ctx *c, u32 i) { return c->q[i]; }
u64 orc_psi(const orc_ctx *c, u32 i) { return c->psi[i]; }

/* ---------------------------------------------------------------- NTT */
/* SEAL util::ntt_negacyclic_harvey (non-lazy): Cooley-Tukey, natural-order input,
 * bit-reversed-order output a^[k] = a(psi^{2 brev(k)+1}), canonical residues.
 * ntt.cpp / dwthandler.h; SURVEY 9.2.  Harvey lazy butterflies keep values in [0,4q). */
void orc_ntt_fwd(const orc_ctx *c, u32 mi, u64 *a)
{
    const u32 n = c->n;
    const u64 q = c->q[mi], q2 = q << 1;
    const u64 *w = c->w[mi], *wsh = c->wsh[mi];
    for (u32 m = 1, gap = n >> 1; m < n; m <<= 1, gap >>= 1) {
        for (u32 i = 0; i < m; ++i) {
            const u64 W = w[m + i], Wsh = wsh[m + i];
            u64 *x = a + 2 * i * gap, *y = x + gap;
            for (u32 j = 0; j < gap; ++j) {
                u64 X = x[j];
                X -= (X >= q2) ? q2 : 0;
                u64 T = mul_shoup_lazy(y[j], W, Wsh, q);
                x[j] = X + T;
                y[j] = X + q2 - T;
            }
        }
    }
    for (u32 j = 0; j < n; ++j) {
        u64 v = a[j];
        v -= (v >= q2) ? q2 : 0;
        v -= (v >= q) ? q : 0;
        a[j] = v;
    }
}

/* SEAL util::inverse_ntt_negacyclic_harvey (non-lazy): Gentleman-Sande, exact inverse
 * of orc_ntt_fwd with N^{-1} folded in, canonical output.  SURVEY 9.2. */
void orc_ntt_inv(const orc_ctx *c, u32 mi, u64 *a)
{
    const u32 n = c->n;
    const u64 q = c->q[mi], q2 = q << 1;
    const u64 *w = c->iw[mi], *wsh = c->iwsh[mi];
    for (u32 m = n >> 1, gap = 1; m >= 1; m >>= 1, gap <<= 1) {
        for (u32 i = 0; i < m; ++i) {
            const u64 W = w[m + i], Wsh = wsh[m + i];
            u64 *x = a + 2 * i * gap, *y = x + gap;
            for (u32 j = 0; j < gap; ++j) {
                u64 X = x[j], Y = y[j]; /* both in [0,2q) */
                u64 S = X + Y;
                S -= (S >= q2) ? q2 : 0;
                x[j] = S;
                y[j] = mul_shoup_lazy(X + q2 - Y, W, Wsh, q);
            }
        }
    }
    const u64 ni = c->ninv[mi], nish = c->ninv_sh[mi];
    for (u32 j = 0; j < n; ++j) {
        u64 v = mul_shoup_lazy(a[j], ni, nish, q);
        v -= (v >= q) ? q : 0;
        a[j] = v;
    }
}

/* ---------------------------------------------------------------- element-wise */
#define POLY(p, L, n, k, i) ((p) + ((size_t)(k) * (L) + (i)) * (n))

/* Evaluator::negate -- reached from he_operators.cpp:16,26 */
void orc_negate(const orc_ctx *c, u32 L, const u64 *a, u32 size, u64 *out)
{
    const u32 n = c->n;
    for (u32 k = 0; k < size; ++k)
        for (u32 i = 0; i < L; ++i) {
            const u64 q = c->q[i];
            const u64 *x = POLY(a, L, n, k, i);
            u64 *o = POLY(out, L, n, k, i);
            for (u32 j = 0; j < n; ++j) o[j] = negmod(x[j], q);
        }
}

/* Evaluator::add -- he_operators.cpp:35,45.  Sizes may differ: the extra polynomials of
 * the larger operand are copied. out has max(sa,sb) polynomials. */
void orc_add(const orc_ctx *c, u32 L, const u64 *a, u32 sa, const u64 *b, u32 sb, u64 *out)
{
    const u32 n = c->n;
    u32 mn = sa < sb ? sa : sb, mx = sa < sb ? sb : sa;
    for (u32 k = 0; k < mx; ++k)
        for (u32 i = 0; i < L; ++i) {
            const u64 q = c->q[i];
            u64 *o = POLY(out, L, n, k, i);
            if (k < mn) {
                const u64 *x = POLY(a, L, n, k, i), *y = POLY(b, L, n, k, i);
                for (u32 j = 0; j < n; ++j) o[j] = addmod(x[j], y[j], q);
            } else {
                const u64 *x = sa > sb ? POLY(a, L, n, k, i) : POLY(b, L, n, k, i);
                memmove(o, x, sizeof(u64) * n);
            }
        }
}

/* Evaluator::sub -- he_operators.cpp:73,83.  Extra polynomials of a larger subtrahend
 * are negated. */
void orc_sub(const orc_ctx *c, u32 L, const u64 *a, u32 sa, const u64 *b, u32 sb, u64 *out)
{
    const u32 n = c->n;
    u32 mn = sa < sb ? sa : sb, mx = sa < sb ? sb : sa;
    for (u32 k = 0; k < mx; ++k)
        for (u32 i = 0; i < L; ++i) {
            const u64 q = c->q[i];
            u64 *o = POLY(out, L, n, k, i);
            if (k < mn) {
                const u64 *x = POLY(a, L, n, k, i), *y = POLY(b, L, n, k, i);
                for (u32 j = 0; j < n; ++j) o[j] = submod(x[j], y[j], q);
            } else if (sa > sb) {
                memmove(o, POLY(a, L, n, k, i), sizeof(u64) * n);
            } else {
                const u64 *y = POLY(b, L, n, k, i);
                for (u32 j = 0; j < n; ++j) o[j] = negmod(y[j], q);
            }
        }
}

/* Evaluator::add_plain / sub_plain (CKKS, NTT form): plaintext added to c0 only --
 * he_operators.cpp:54,64,92,102 */
void orc_add_plain(const orc_ctx *c, u32 L, const u64 *ct, u32 size, const u64 *pt, u64 *out)
{
    const u32 n = c->n;
    if (out != ct) memmove(out, ct, sizeof(u64) * (size_t)size * L * n);
    for (u32 i = 0; i < L; ++i) {
        const u64 q = c->q[i];
        u64 *o = POLY(out, L, n, 0, i);
        const u64 *p = pt + (size_t)i * n;
        for (u32 j = 0; j < n; ++j) o[j] = addmod(o[j], p[j], q);
    }
}
void orc_sub_plain(const orc_ctx *c, u32 L, const u64 *ct, u32 size, const u64 *pt, u64 *out)
{
    const u32 n = c->n;
    if (out != ct) memmove(out, ct, sizeof(u64) * (size_t)size * L * n);
    for (u32 i = 0; i < L; ++i) {
        const u64 q = c->q[i];
        u64 *o = POLY(out, L, n, 0, i);
        const u64 *p = pt + (size_t)i * n;
        for (u32 j = 0; j < n; ++j) o[j] = submod(o[j], p[j], q);
    }
}

/* Evaluator::multiply_plain_ntt -- he_operators.cpp:130,140; he_util.h:35,43 */
void orc_multiply_plain(const orc_ctx *c, u32 L, const u64 *ct, u32 size, const u64 *pt, u64 *out)
{
    const u32 n = c->n;
    for (u32 k = 0; k < size; ++k)
        for (u32 i = 0; i < L; ++i) {
            const u64 q = c->q[i];
            const u64 *x = POLY(ct, L, n, k, i), *p = pt + (size_t)i * n;
            u64 *o = POLY(out, L, n, k, i);
            for (u32 j = 0; j < n; ++j) o[j] = MULMOD(c, i, x[j], p[j]);
            (void)q;
        }
}

/* Evaluator::ckks_multiply: dyadic convolution d_k = sum_{i+j=k} a_i b_j --
 * he_operators.cpp:111,121.  out has sa+sb-1 polynomials and must not alias. */
void orc_multiply(const orc_ctx *c, u32 L, const u64 *a, u32 sa, const u64 *b, u32 sb, u64 *out)
{
    const u32 n = c->n;
    u32 so = sa + sb - 1;
    for (u32 k = 0; k < so; ++k)
        for (u32 i = 0; i < L; ++i) {
            const u64 q = c->q[i];
            u64 *o = POLY(out, L, n, k, i);
            for (u32 j = 0; j < n; ++j) {
                u64 acc = 0;
                for (u32 ia = 0; ia < sa; ++ia) {
                    if (k < ia || k - ia >= sb) continue;
                    acc = addmod(acc, MULMOD(c, i, POLY(a, L, n, ia, i)[j], POLY(b, L, n, k - ia, i)[j]), q);
                }
                o[j] = acc;
            }
        }
}

/* Evaluator::ckks_square -- he_linalg.cpp:647.  Same function as multiply(a, a). */
void orc_square(const orc_ctx *c, u32 L, const u64 *a, u32 sa, u64 *out) { orc_multiply(c, L, a, sa, a, sa, out); }

/* ---------------------------------------------------------------- rescale */
/* RNSTool::divide_and_round_q_last_ntt_inplace + limb drop --
 * he_operators.cpp:168,178; he_util.h:36,44.  ct [size][L][N] -> out [size][L-1][N]. */
void orc_rescale(const orc_ctx *c, u32 L, const u64 *ct, u32 size, u64 *out)
{
    const u32 n = c->n;
    const u64 ql = c->q[L - 1], half = ql >> 1;
    u64 *t = (u64 *)malloc(sizeof(u64) * n), *d = (u64 *)malloc(sizeof(u64) * n);
    for (u32 k = 0; k < size; ++k) {
        memcpy(t, POLY(ct, L, n, k, L - 1), sizeof(u64) * n);
        orc_ntt_inv(c, L - 1, t);
        for (u32 j = 0; j < n; ++j) t[j] = addmod(t[j], half, ql);
        for (u32 i = 0; i + 1 < L; ++i) {
            const u64 q = c->q[i];
            const u64 hq = half % q, inv = invmod(ql % q, q);
            for (u32 j = 0; j < n; ++j) d[j] = submod(REDUCE64(c, i, t[j]), hq, q);
            orc_ntt_fwd(c, i, d);
            const u64 *x = POLY(ct, L, n, k, i);
            u64 *o = POLY(out, L - 1, n, k, i);
            for (u32 j = 0; j < n; ++j) o[j] = MULMOD(c, i, submod(x[j], d[j], q), inv);
        }
    }
    free(t);
    free(d);
}

/* Evaluator::mod_switch_drop_to_next (CKKS): drop the last limb -- he_operators.cpp:187,197 */
void orc_mod_switch(const orc_ctx *c, u32 L, const u64 *ct, u32 size, u64 *out)
{
    const u32 n = c->n;
    for (u32 k = 0; k < size; ++k)
        for (u32 i = 0; i + 1 < L; ++i) memmove(POLY(out, L - 1, n, k, i), POLY(ct, L, n, k, i), sizeof(u64) * n);
}

/* ---------------------------------------------------------------- Galois */
/* GaloisTool::get_elt_from_step -- galois.cpp; SURVEY 9.4 */
u32 orc_galois_elt_from_step(u32 n, int step)
{
    u32 m = 2 * n;
    if (step == 0) return m - 1;
    u32 pos = (u32)(step < 0 ? -step : step);
    u32 s = step < 0 ? (n >> 1) - pos : pos;
    u64 e = 1;
    for (u32 i = 0; i < s; ++i) e = (e * 3) & (m - 1);
    return (u32)e;
}

/* util::naf -- numth.h; SURVEY 9.4 (LSB-first signed digits) */
int orc_naf(int value, int *out)
{
    int cnt = 0;
    int sign = value < 0;
    if (sign) value = -value;
    for (int i = 0; value; ++i) {
        int zi = (value & 1) ? 2 - (value & 3) : 0;
        value = (value - zi) >> 1;
        if (zi) out[cnt++] = (sign ? -zi : zi) * (1 << i);
    }
    return cnt;
}

/* GaloisTool::generate_table_ntt -- galois.cpp; SURVEY 9.4:
 * table[i] = brev(((elt*(2 brev(i)+1)) >> 1) & (N-1)); result[i] = operand[table[i]] */
void orc_galois_table(u32 n, u32 elt, u32 *table)
{
    u32 logn = 0;
    while ((1u << logn) < n) logn++;
    for (u32 i = 0; i < n; ++i) {
        u64 r = 2ull * brev(i, logn) + 1;
        u64 raw = ((u64)elt * r) >> 1;
        table[i] = brev((u32)(raw & (n - 1)), logn);
    }
}

/* GaloisTool::apply_galois_ntt over `limbs` limbs; out must not alias in */
void orc_apply_galois_ntt(const orc_ctx *c, u32 limbs, u32 elt, const u64 *in, u64 *out)
{
    const u32 n = c->n;
    u32 *tab = (u32 *)malloc(sizeof(u32) * n);
    orc_galois_table(n, elt, tab);
    for (u32 i = 0; i < limbs; ++i)
        for (u32 j = 0; j < n; ++j) out[(size_t)i * n + j] = in[(size_t)i * n + tab[j]];
    free(tab);
}

/* ---------------------------------------------------------------- key switching */
/* digit decomposition of switch_key_inplace:
 * ext[j][I] = NTT_{m_I}( INTT_{q_j}(target_j) mod m_I ), I in [0,L] (I = L is the special prime);
 * ext[j][j] = target_j itself.  ext layout [L][L+1][N]. */
static void ks_decompose(const orc_ctx *c, u32 L, const u64 *target, u64 *ext)
{
    const u32 n = c->n, K = c->K;
    u64 *coef = (u64 *)malloc(sizeof(u64) * n);
    for (u32 J = 0; J < L; ++J) {
        memcpy(coef, target + (size_t)J * n, sizeof(u64) * n);
        orc_ntt_inv(c, J, coef);
        for (u32 I = 0; I <= L; ++I) {
            u64 *e = ext + ((size_t)J * (L + 1) + I) * n;
            if (I == J) {
                memcpy(e, target + (size_t)J * n, sizeof(u64) * n);
                continue;
            }
            const u32 ki = (I == L) ? K - 1 : I;
            const u64 m = c->q[ki];
            if (c->q[J] <= m)
                memcpy(e, coef, sizeof(u64) * n);
            else
                for (u32 x = 0; x < n; ++x) e[x] = REDUCE64(c, ki, coef[x]);
            (void)m;
            orc_ntt_fwd(c, ki, e);
        }
    }
    free(coef);
}

/* key inner product (SURVEY 9.6 step 2, 128-bit lazy sums, one reduction):
 * acc[comp][I] = sum_j ext[j][I][tab[x]] * key[j][comp][idx(I)][x] mod m_I; tab = Galois gather
 * table applied to the lifted digits (hoisted rotation) or NULL.  acc layout [2][L+1][N]. */
static void ks_inner(const orc_ctx *c, u32 L, const u64 *ext, const u32 *tab, const u64 *key, u64 *acc)
{
    const u32 n = c->n, K = c->K;
    u128 *lazy = (u128 *)malloc(sizeof(u128) * (size_t)2 * n);
    for (u32 I = 0; I <= L; ++I) {
        const u32 ki = (I == L) ? K - 1 : I;
        const u64 m = c->q[ki];
        memset(lazy, 0, sizeof(u128) * (size_t)2 * n);
        for (u32 J = 0; J < L; ++J) {
            const u64 *op = ext + ((size_t)J * (L + 1) + I) * n;
            for (u32 comp = 0; comp < 2; ++comp) {
                const u64 *kp = key + (((size_t)J * 2 + comp) * K + ki) * n;
                u128 *l = lazy + (size_t)comp * n;
                if (tab)
                    for (u32 x = 0; x < n; ++x) l[x] += (u128)op[tab[x]] * kp[x];
                else
                    for (u32 x = 0; x < n; ++x) l[x] += (u128)op[x] * kp[x];
            }
        }
        for (u32 comp = 0; comp < 2; ++comp) {
            u64 *a = acc + ((size_t)comp * (L + 1) + I) * n;
            const u128 *l = lazy + (size_t)comp * n;
            for (u32 x = 0; x < n; ++x) a[x] = barrett128(l[x], m, c->r0[ki], c->r1[ki]);
        }
    }
    free(lazy);
}

/* mod-down by P with rounding: out[comp][i] = base[comp][i] +
 * (acc[comp][i] - NTT_{q_i}((t mod q_i) - (floor(P/2) mod q_i))) * P^-1, t = INTT_P(acc[comp][L]) + floor(P/2).
 * acc [2][L+1][N] is destroyed; base/out [2][L][N] may alias. */
static void ks_mod_down_add(const orc_ctx *c, u32 L, u64 *acc, const u64 *base, u64 *out)
{
    const u32 n = c->n, K = c->K;
    const u64 P = c->q[K - 1], halfP = P >> 1;
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
    for (u32 comp = 0; comp < 2; ++comp) {
        u64 *t = acc + ((size_t)comp * (L + 1) + L) * n;
        orc_ntt_inv(c, K - 1, t);
        for (u32 x = 0; x < n; ++x) t[x] = addmod(t[x], halfP, P);
        for (u32 i = 0; i < L; ++i) {
            const u64 q = c->q[i];
            const u64 hq = halfP % q, pinv = invmod(P % q, q);
            for (u32 x = 0; x < n; ++x) tmp[x] = submod(REDUCE64(c, i, t[x]), hq, q);
            orc_ntt_fwd(c, i, tmp);
            const u64 *a = acc + ((size_t)comp * (L + 1) + i) * n;
            const u64 *b = POLY(base, L, n, comp, i);
            u64 *o = POLY(out, L, n, comp, i);
            for (u32 x = 0; x < n; ++x) o[x] = addmod(b[x], MULMOD(c, i, submod(a[x], tmp[x], q), pinv), q);
        }
    }
    free(tmp);
}

/* Evaluator::switch_key_inplace (CKKS) -- evaluator.cpp; SURVEY 9.6.
 * ct [2][L][N] (in/out), target [L][N] NTT form, key [Lmax][2][K][N], Lmax = K-1. */
void orc_switch_key(const orc_ctx *c, u32 L, u64 *ct, const u64 *target, const u64 *key)
{
    const u32 n = c->n;
    u64 *ext = (u64 *)malloc(sizeof(u64) * (size_t)L * (L + 1) * n);
    u64 *acc = (u64 *)malloc(sizeof(u64) * (size_t)2 * (L + 1) * n);
    ks_decompose(c, L, target, ext);
    ks_inner(c, L, ext, NULL, key, acc);
    ks_mod_down_add(c, L, acc, ct, ct);
    free(ext);
    free(acc);
}

/* Evaluator::relinearize_internal (size 3 -> 2) -- he_operators.cpp:149,159 */
void orc_relinearize(const orc_ctx *c, u32 L, const u64 *ct3, const u64 *rk, u64 *out)
{
    const u32 n = c->n;
    u64 *target = (u64 *)malloc(sizeof(u64) * (size_t)L * n);
    memcpy(target, ct3 + (size_t)2 * L * n, sizeof(u64) * (size_t)L * n);
    memmove(out, ct3, sizeof(u64) * (size_t)2 * L * n);
    orc_switch_key(c, L, out, target, rk);
    free(target);
}

/* Evaluator::apply_galois_inplace (CKKS, size 2) -- SURVEY 9.4:
 * c0 <- pi(c0); t <- pi(c1); c1 <- 0; switch_key(ct, t, key[elt]) */
void orc_apply_galois(const orc_ctx *c, u32 L, const u64 *ct, u32 elt, const u64 *key, u64 *out)
{
    const u32 n = c->n;
    u64 *res = (u64 *)calloc((size_t)2 * L * n, sizeof(u64));
    u64 *t = (u64 *)malloc(sizeof(u64) * (size_t)L * n);
    orc_apply_galois_ntt(c, L, elt, ct, res);
    orc_apply_galois_ntt(c, L, elt, ct + (size_t)L * n, t);
    orc_switch_key(c, L, res, t, key);
    memcpy(out, res, sizeof(u64) * (size_t)2 * L * n);
    free(res);
    free(t);
}

static const u64 *find_key(u32 elt, u32 n_keys, const u32 *elts, const u64 *const *keys)
{
    for (u32 i = 0; i < n_keys; ++i)
        if (elts[i] == elt) return keys[i];
    return NULL;
}

/* Evaluator::rotate_internal -- evaluator.cpp; SURVEY 9.4.  Reached from
 * he_operators.cpp:206-235 and he_linalg.cpp:595,609,622,636. */
int orc_rotate(const orc_ctx *c, u32 L, const u64 *ct, int steps, u32 n_keys, const u32 *elts,
               const u64 *const *keys, u64 *out)
{
    const u32 n = c->n;
    size_t bytes = sizeof(u64) * (size_t)2 * L * n;
    if (out != ct) memmove(out, ct, bytes);
    if (steps == 0) return 0;
    const u64 *k = find_key(orc_galois_elt_from_step(n, steps), n_keys, elts, keys);
    if (k) {
        orc_apply_galois(c, L, out, orc_galois_elt_from_step(n, steps), k, out);
        return 1;
    }
    int terms[40];
    int cnt = orc_naf(steps, terms);
    if (cnt == 1) return -1; /* "Galois key not present" */
    int total = 0;
    for (int i = 0; i < cnt; ++i) {
        int a = terms[i] < 0 ? -terms[i] : terms[i];
        if ((u32)a == (n >> 1)) continue;
        int r = orc_rotate(c, L, out, terms[i], n_keys, elts, keys, out);
        if (r < 0) return -1;
        total += r;
    }
    return total;
}

/* ---------------------------------------------------------------- client side */
static inline u64 splitmix(u64 *s)
{
    u64 z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline u64 uniform_mod(u64 *s, u64 q)
{
    /* rejection sampling as SEAL sample_poly_uniform */
    u64 max_multiple = UINT64_MAX - (UINT64_MAX % q) - 1;
    u64 r;
    do r = splitmix(s); while (r >= max_multiple);
    return r % q;
}
/* SEAL sample_poly_cbd: centered binomial, 21 bits each side (sigma ~ 3.24) */
static inline int cbd(u64 *s)
{
    u64 r = splitmix(s);
    return __builtin_popcountll(r & 0x1FFFFF) - __builtin_popcountll((r >> 21) & 0x1FFFFF);
}

/* SecretKey: ternary coefficients, stored NTT form over all K primes ([K][N]) */
void orc_sample_secret(const orc_ctx *c, u64 seed, u64 *s)
{
    const u32 n = c->n, K = c->K;
    u64 st = seed;
    for (u32 j = 0; j < n; ++j) {
        u64 r = uniform_mod(&st, 3); /* 0,1,2 -> 0, 1, -1 */
        for (u32 i = 0; i < K; ++i) s[(size_t)i * n + j] = r == 0 ? 0 : (r == 1 ? 1 : c->q[i] - 1);
    }
    for (u32 i = 0; i < K; ++i) orc_ntt_fwd(c, i, s + (size_t)i * n);
}

/* encrypt_zero_symmetric at `limbs` key-level limbs: (c0,c1) = (-(a s + e), a), NTT form.
 * limb_map[i] = index into ctx moduli for limb i. ct layout [2][limbs][N]. */
static void encrypt_zero(const orc_ctx *c, u64 *st, u32 limbs, const u32 *limb_map, const u64 *s, u64 *ct)
{
    const u32 n = c->n;
    int *e = (int *)malloc(sizeof(int) * n);
    u64 *en = (u64 *)malloc(sizeof(u64) * n);
    for (u32 j = 0; j < n; ++j) e[j] = cbd(st);
    for (u32 i = 0; i < limbs; ++i) {
        const u32 mi = limb_map[i];
        const u64 q = c->q[mi];
        u64 *c0 = ct + (size_t)i * n, *c1 = ct + ((size_t)limbs + i) * n;
        for (u32 j = 0; j < n; ++j) c1[j] = uniform_mod(st, q);
        for (u32 j = 0; j < n; ++j) en[j] = e[j] >= 0 ? (u64)e[j] : q - (u64)(-e[j]);
        orc_ntt_fwd(c, mi, en);
        const u64 *sp = s + (size_t)mi * n;
        for (u32 j = 0; j < n; ++j) c0[j] = negmod(addmod(mulmod(c1[j], sp[j], q), en[j], q), q);
    }
    free(e);
    free(en);
}

/* KeyGenerator::generate_one_kswitch_key -- keygenerator.cpp; SURVEY 9.5.
 * new_key [K][N] NTT form; out [Lmax][2][K][N]. */
void orc_gen_kswitch_key(const orc_ctx *c, u64 seed, const u64 *s, const u64 *new_key, u64 *out)
{
    const u32 n = c->n, K = c->K, Lmax = K - 1;
    u64 st = seed;
    u32 map[ORC_MAXK];
    for (u32 i = 0; i < K; ++i) map[i] = i;
    for (u32 j = 0; j < Lmax; ++j) {
        u64 *kj = out + (size_t)j * 2 * K * n;
        encrypt_zero(c, &st, K, map, s, kj);
        const u64 q = c->q[j];
        const u64 factor = c->q[K - 1] % q;
        u64 *dst = kj + (size_t)j * n; /* component 0, limb j */
        const u64 *nk = new_key + (size_t)j * n;
        for (u32 x = 0; x < n; ++x) dst[x] = addmod(dst[x], mulmod(nk[x], factor, q), q);
    }
}

/* KeyGenerator::create_relin_keys: new key = s^2 */
void orc_gen_relin_key(const orc_ctx *c, u64 seed, const u64 *s, u64 *out)
{
    const u32 n = c->n, K = c->K;
    u64 *s2 = (u64 *)malloc(sizeof(u64) * (size_t)K * n);
    for (u32 i = 0; i < K; ++i)
        for (u32 j = 0; j < n; ++j) s2[(size_t)i * n + j] = mulmod(s[(size_t)i * n + j], s[(size_t)i * n + j], c->q[i]);
    orc_gen_kswitch_key(c, seed, s, s2, out);
    free(s2);
}

/* KeyGenerator::create_galois_keys (one element): new key = pi_elt(s) */
void orc_gen_galois_key(const orc_ctx *c, u64 seed, const u64 *s, u32 elt, u64 *out)
{
    const u32 n = c->n, K = c->K;
    u64 *rs = (u64 *)malloc(sizeof(u64) * (size_t)K * n);
    orc_apply_galois_ntt(c, K, elt, s, rs);
    orc_gen_kswitch_key(c, seed, s, rs, out);
    free(rs);
}

/* Encryptor::encrypt_symmetric at level L: ct = (-(a s + e) + m, a) */
void orc_encrypt_symmetric(const orc_ctx *c, u64 seed, u32 L, const u64 *s, const u64 *plain, u64 *ct)
{
    const u32 n = c->n;
    u64 st = seed;
    u32 map[ORC_MAXK];
    for (u32 i = 0; i < ORC_MAXK; ++i) map[i] = i;
    encrypt_zero(c, &st, L, map, s, ct);
    for (u32 i = 0; i < L; ++i)
        for (u32 j = 0; j < n; ++j) ct[(size_t)i * n + j] = addmod(ct[(size_t)i * n + j], plain[(size_t)i * n + j], c->q[i]);
}

/* Decryptor::ckks_decrypt: m = sum_k c_k s^k (NTT form result) */
void orc_decrypt(const orc_ctx *c, u32 L, const u64 *ct, u32 size, const u64 *s, u64 *plain)
{
    const u32 n = c->n;
    for (u32 i = 0; i < L; ++i) {
        const u64 q = c->q[i];
        const u64 *sp = s + (size_t)i * n;
        u64 *o = plain + (size_t)i * n;
        for (u32 j = 0; j < n; ++j) {
            u64 acc = 0;
            for (u32 k = size; k-- > 0;) acc = addmod(mulmod(acc, sp[j], q), POLY(ct, L, n, k, i)[j], q);
            o[j] = acc;
        }
    }
}

/* ---------------------------------------------------------------- composites */
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* BSGS diagonal matvec, the restatement of the CUDA composite hegpu_matvec_bsgs (which
 * replaces the per-diagonal loop of he_linalg.cpp:977-1003 / he_fft.cpp:178-203 for
 * plaintext diagonals).  Built only from the primitives above, in this order:
 *   baby_b = apply_galois(ct, elt(b))                         b = 1..n1-1  (baby_0 = ct)
 *   inner_g = sum_b multiply_plain(baby_b, pt[g*n1+b])        accumulation order b = 0..n1-1
 *   acc = inner_0 + sum_{g>=1} apply_galois(inner_g, elt(g*n1))   order g = 1..n2-1
 *   out = rescale(acc)
 */
void orc_matvec_bsgs(const orc_ctx *c, u32 L, u32 B, const u64 *cts, u32 n1, u32 n2, const u64 *pts,
                     const u64 *const *baby_keys, const u64 *const *giant_keys, u64 *out, int threads)
{
    const u32 n = c->n;
    const size_t ctw = (size_t)2 * L * n, ptw = (size_t)L * n;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
#endif
    for (u32 b = 0; b < B; ++b) {
        const u64 *ct = cts + (size_t)b * ctw;
        u64 *baby = (u64 *)malloc(sizeof(u64) * ctw * n1);
        u64 *inner = (u64 *)malloc(sizeof(u64) * ctw);
        u64 *prod = (u64 *)malloc(sizeof(u64) * ctw);
        u64 *acc = (u64 *)malloc(sizeof(u64) * ctw);
        memcpy(baby, ct, sizeof(u64) * ctw);
        for (u32 i = 1; i < n1; ++i) orc_apply_galois(c, L, ct, orc_galois_elt_from_step(n, (int)i), baby_keys[i], baby + (size_t)i * ctw);
        for (u32 g = 0; g < n2; ++g) {
            for (u32 i = 0; i < n1; ++i) {
                const u64 *pt = pts + (size_t)(g * n1 + i) * ptw;
                if (i == 0) {
                    orc_multiply_plain(c, L, baby, 2, pt, inner);
                } else {
                    orc_multiply_plain(c, L, baby + (size_t)i * ctw, 2, pt, prod);
                    orc_add(c, L, inner, 2, prod, 2, inner);
                }
            }
            if (g == 0) {
                memcpy(acc, inner, sizeof(u64) * ctw);
            } else {
                orc_apply_galois(c, L, inner, orc_galois_elt_from_step(n, (int)(g * n1)), giant_keys[g], prod);
                orc_add(c, L, acc, 2, prod, 2, acc);
            }
        }
        orc_rescale(c, L, acc, 2, out + (size_t)b * 2 * (L - 1) * n);
        free(baby);
        free(inner);
        free(prod);
        free(acc);
    }
}

/* General BSGS matvec = the restatement of hegpu_matvec_bsgs_range.  flags: 1 = hoisted baby steps,
 * 2 = lazy giant steps (one mod-down), 4 = final rescale.  This call owns global giant steps
 * g_first .. g_first+n2-1 (diagonal sharding); giant_keys[g'] is the key of rotation
 * (g_first+g')*n1 and is unused for global step 0.
 *  - hoisted baby steps: the digits of c1 are decomposed once (ks_decompose, no permutation) and
 *    every rotation applies its Galois permutation to the lifted digits:
 *       baby_k = mod_down( ks_inner(pi_k(ext), key_k) ) + (pi_k(c0), 0)
 *  - lazy giant steps: acc = sum_g ks_inner(decompose(pi_g(inner_g.c1)), key_g) in the extended
 *    basis, ONE mod-down, base = ([inner_0.c0] + sum_g pi_g(inner_g.c0), [inner_0.c1]).
 * Same function as the chain of primitives up to key-switch noise; different bits.
 * out: [B][2][L-1][N] with rescale, [B][2][L][N] without. */
void orc_matvec_bsgs_ex(const orc_ctx *c, u32 L, u32 B, const u64 *cts, u32 n1, u32 n2, u32 g_first, const u64 *pts,
                        const u64 *const *baby_keys, const u64 *const *giant_keys, int flags, u64 *out, int threads)
{
    const u32 n = c->n, K = c->K;
    const int hoist = flags & 1, lazy = flags & 2, rescale = flags & 4;
    const size_t ctw = (size_t)2 * L * n, ptw = (size_t)L * n, accw = (size_t)2 * (L + 1) * n;
    u32 *tabs = (u32 *)malloc(sizeof(u32) * (size_t)(n1 + n2) * n);
    for (u32 k = 1; k < n1; ++k) orc_galois_table(n, orc_galois_elt_from_step(n, (int)k), tabs + (size_t)k * n);
    for (u32 g = 0; g < n2; ++g)
        if (g_first + g) orc_galois_table(n, orc_galois_elt_from_step(n, (int)((g_first + g) * n1)), tabs + (size_t)(n1 + g) * n);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
#endif
    for (u32 b = 0; b < B; ++b) {
        const u64 *ct = cts + (size_t)b * ctw;
        u64 *ext = (u64 *)malloc(sizeof(u64) * (size_t)L * (L + 1) * n);
        u64 *acc = (u64 *)malloc(sizeof(u64) * accw);
        u64 *accsum = (u64 *)calloc(accw, sizeof(u64));
        u64 *baby = (u64 *)malloc(sizeof(u64) * ctw * n1);
        u64 *inner = (u64 *)malloc(sizeof(u64) * ctw * n2);
        u64 *prod = (u64 *)malloc(sizeof(u64) * ctw);
        u64 *base = (u64 *)calloc(ctw, sizeof(u64));
        u64 *res = (u64 *)calloc(ctw, sizeof(u64));
        memcpy(baby, ct, sizeof(u64) * ctw);
        if (hoist) ks_decompose(c, L, ct + (size_t)L * n, ext);
        for (u32 k = 1; k < n1; ++k) {
            if (!hoist) {
                orc_apply_galois(c, L, ct, orc_galois_elt_from_step(n, (int)k), baby_keys[k], baby + (size_t)k * ctw);
                continue;
            }
            const u32 *tab = tabs + (size_t)k * n;
            ks_inner(c, L, ext, tab, baby_keys[k], acc);
            memset(base, 0, sizeof(u64) * ctw);
            for (u32 i = 0; i < L; ++i)
                for (u32 x = 0; x < n; ++x) base[(size_t)i * n + x] = ct[(size_t)i * n + tab[x]];
            ks_mod_down_add(c, L, acc, base, baby + (size_t)k * ctw);
        }
        for (u32 g = 0; g < n2; ++g) {
            u64 *ig = inner + (size_t)g * ctw;
            for (u32 i = 0; i < n1; ++i) {
                const u64 *pt = pts + (size_t)(g * n1 + i) * ptw;
                if (i == 0) {
                    orc_multiply_plain(c, L, baby, 2, pt, ig);
                } else {
                    orc_multiply_plain(c, L, baby + (size_t)i * ctw, 2, pt, prod);
                    orc_add(c, L, ig, 2, prod, 2, ig);
                }
            }
        }
        int rotated = 0;
        memset(base, 0, sizeof(u64) * ctw);
        for (u32 g = 0; g < n2; ++g) {
            const u64 *ig = inner + (size_t)g * ctw;
            if (g_first + g == 0) { /* unrotated term */
                if (lazy) orc_add(c, L, base, 2, ig, 2, base);
                else orc_add(c, L, res, 2, ig, 2, res);
                continue;
            }
            if (!lazy) {
                orc_apply_galois(c, L, ig, orc_galois_elt_from_step(n, (int)((g_first + g) * n1)), giant_keys[g], prod);
                orc_add(c, L, res, 2, prod, 2, res);
                continue;
            }
            rotated = 1;
            const u32 *tab = tabs + (size_t)(n1 + g) * n;
            for (u32 i = 0; i < L; ++i) /* target = pi_g(c1) */
                for (u32 x = 0; x < n; ++x) prod[(size_t)i * n + x] = ig[(size_t)(L + i) * n + tab[x]];
            ks_decompose(c, L, prod, ext);
            ks_inner(c, L, ext, NULL, giant_keys[g], acc);
            for (u32 comp = 0; comp < 2; ++comp)
                for (u32 I = 0; I <= L; ++I) {
                    const u64 m = c->q[I == L ? K - 1 : I];
                    u64 *s = accsum + ((size_t)comp * (L + 1) + I) * n;
                    const u64 *a = acc + ((size_t)comp * (L + 1) + I) * n;
                    for (u32 x = 0; x < n; ++x) s[x] = addmod(s[x], a[x], m);
                }
            for (u32 i = 0; i < L; ++i) {
                const u64 q = c->q[i];
                for (u32 x = 0; x < n; ++x) base[(size_t)i * n + x] = addmod(base[(size_t)i * n + x], ig[(size_t)i * n + tab[x]], q);
            }
        }
        if (lazy) {
            if (rotated) ks_mod_down_add(c, L, accsum, base, res);
            else memcpy(res, base, sizeof(u64) * ctw);
        }
        if (rescale) orc_rescale(c, L, res, 2, out + (size_t)b * 2 * (L - 1) * n);
        else memcpy(out + (size_t)b * ctw, res, sizeof(u64) * ctw);
        free(ext); free(acc); free(accsum); free(baby); free(inner); free(prod); free(base); free(res);
    }
    free(tabs);
}

/* the HEGPU_MATVEC_HOIST | HEGPU_MATVEC_LAZY | HEGPU_MATVEC_RESCALE mode */
void orc_matvec_bsgs_fast(const orc_ctx *c, u32 L, u32 B, const u64 *cts, u32 n1, u32 n2, const u64 *pts,
                          const u64 *const *baby_keys, const u64 *const *giant_keys, u64 *out, int threads)
{
    orc_matvec_bsgs_ex(c, L, B, cts, n1, n2, 0, pts, baby_keys, giant_keys, 1 | 2 | 4, out, threads);
}

/* mod-down by P with rounding of ONE component: out[i] = (acc[i] - NTT_{q_i}((t mod q_i) - (floor(P/2) mod q_i))) * P^-1,
 * t = INTT_P(acc[L]) + floor(P/2).  acc [L+1][N] is destroyed; out [L][N]. */
static void mod_down_one(const orc_ctx *c, u32 L, u64 *acc, u64 *out)
{
    const u32 n = c->n, K = c->K;
    const u64 P = c->q[K - 1], halfP = P >> 1;
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
    u64 *t = acc + (size_t)L * n;
    orc_ntt_inv(c, K - 1, t);
    for (u32 x = 0; x < n; ++x) t[x] = addmod(t[x], halfP, P);
    for (u32 i = 0; i < L; ++i) {
        const u64 q = c->q[i];
        const u64 hq = halfP % q, pinv = invmod(P % q, q);
        for (u32 x = 0; x < n; ++x) tmp[x] = submod(REDUCE64(c, i, t[x]), hq, q);
        orc_ntt_fwd(c, i, tmp);
        const u64 *a = acc + (size_t)i * n;
        u64 *o = out + (size_t)i * n;
        for (u32 x = 0; x < n; ++x) o[x] = MULMOD(c, i, submod(a[x], tmp[x], q), pinv);
    }
    free(tmp);
}

/* Double-hoisted BSGS matvec = the restatement of hegpu_matvec_bsgs_range with HEGPU_MATVEC_DH
 * (Bossuat et al., "Efficient bootstrapping for approximate HE with non-sparse keys", EUROCRYPT 2021,
 * section 5 -- the same composite, built from SEAL's key-switch steps of SURVEY 9.6):
 *   - the digits of c1 are decomposed once (ks_decompose);
 *   - baby rotation k stays in the extended basis q_0..q_{L-1}, P, scaled by P:
 *       b_k = ( ks_inner(pi_k(ext), key_k)[0] + P*pi_k(c0),  ks_inner(...)[1] ),   b_0 = (P*c0, P*c1)
 *   - inner sums in the extended basis with plaintexts that carry a limb mod P:
 *       u_g = sum_k ptx[g*n1+k] (.) b_k                                (accumulated lazily, one reduction)
 *   - rotated giant step g: only the component that has to be key-switched leaves the extended basis,
 *       v1 = mod_down(u_g[1]);  F += ks_inner(decompose(pi_g(v1)), key_g);  F[0] += pi_g(u_g[0])
 *     (the Galois permutation acts limb-wise, also on the limb mod P); the unrotated step adds u_g to F;
 *   - ONE final mod-down of both components: res = mod_down(F); optional rescale.
 * (n2 - 1) single-component mod-downs + 1 instead of (n1 - 1) + 1 two-component ones.  Same function as the
 * chain of primitives up to key-switch noise; different bits.  ptsx: [n1*n2][L+1][N], limb L = residues mod
 * the special prime (NTT form).  flags: 4 = final rescale.  out as orc_matvec_bsgs_ex. */
void orc_matvec_bsgs_dh(const orc_ctx *c, u32 L, u32 B, const u64 *cts, u32 n1, u32 n2, u32 g_first, const u64 *ptsx,
                        const u64 *const *baby_keys, const u64 *const *giant_keys, int flags, u64 *out, int threads)
{
    const u32 n = c->n, K = c->K;
    const int rescale = flags & 4;
    const size_t ctw = (size_t)2 * L * n, ptw = (size_t)(L + 1) * n, accw = (size_t)2 * (L + 1) * n;
    const u64 P = c->q[K - 1];
    u32 *tabs = (u32 *)malloc(sizeof(u32) * (size_t)(n1 + n2) * n);
    for (u32 k = 1; k < n1; ++k) orc_galois_table(n, orc_galois_elt_from_step(n, (int)k), tabs + (size_t)k * n);
    for (u32 g = 0; g < n2; ++g)
        if (g_first + g) orc_galois_table(n, orc_galois_elt_from_step(n, (int)((g_first + g) * n1)), tabs + (size_t)(n1 + g) * n);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
#endif
    for (u32 b = 0; b < B; ++b) {
        const u64 *ct = cts + (size_t)b * ctw;
        u64 *ext = (u64 *)malloc(sizeof(u64) * (size_t)L * (L + 1) * n);
        u64 *baby = (u64 *)malloc(sizeof(u64) * accw * n1); /* b_k, layout [2][L+1][N] */
        u64 *u = (u64 *)malloc(sizeof(u64) * accw);
        u64 *acc = (u64 *)malloc(sizeof(u64) * accw);
        u64 *F = (u64 *)calloc(accw, sizeof(u64));
        u64 *v1 = (u64 *)malloc(sizeof(u64) * (size_t)L * n);
        u64 *zero = (u64 *)calloc(ctw, sizeof(u64));
        u64 *res = (u64 *)malloc(sizeof(u64) * ctw);
        u64 *tgt = (u64 *)malloc(sizeof(u64) * (size_t)L * n);
        u128 *lazy = (u128 *)malloc(sizeof(u128) * n);
        ks_decompose(c, L, ct + (size_t)L * n, ext);
        for (u32 k = 0; k < n1; ++k) {
            u64 *bk = baby + (size_t)k * accw;
            const u32 *tab = k ? tabs + (size_t)k * n : NULL;
            if (k) ks_inner(c, L, ext, tab, baby_keys[k], bk);
            else memset(bk, 0, sizeof(u64) * accw);
            for (u32 comp = 0; comp < (k ? 1u : 2u); ++comp)
                for (u32 i = 0; i < L; ++i) {
                    const u64 q = c->q[i], pm = P % q;
                    const u64 *src = ct + ((size_t)comp * L + i) * n;
                    u64 *dst = bk + ((size_t)comp * (L + 1) + i) * n;
                    for (u32 x = 0; x < n; ++x) dst[x] = addmod(dst[x], MULMOD(c, i, src[tab ? tab[x] : x], pm), q);
                }
        }
        for (u32 g = 0; g < n2; ++g) {
            for (u32 comp = 0; comp < 2; ++comp)
                for (u32 I = 0; I <= L; ++I) {
                    const u32 ki = (I == L) ? K - 1 : I;
                    memset(lazy, 0, sizeof(u128) * n);
                    for (u32 k = 0; k < n1; ++k) {
                        const u64 *bk = baby + (size_t)k * accw + ((size_t)comp * (L + 1) + I) * n;
                        const u64 *pt = ptsx + (size_t)(g * n1 + k) * ptw + (size_t)I * n;
                        for (u32 x = 0; x < n; ++x) lazy[x] += (u128)bk[x] * pt[x];
                    }
                    u64 *o = u + ((size_t)comp * (L + 1) + I) * n;
                    for (u32 x = 0; x < n; ++x) o[x] = barrett128(lazy[x], c->q[ki], c->r0[ki], c->r1[ki]);
                }
            const u32 *tab = (g_first + g) ? tabs + (size_t)(n1 + g) * n : NULL;
            const u64 *add1 = u + (size_t)(L + 1) * n; /* what is added to F[1] */
            if (tab) {
                mod_down_one(c, L, u + (size_t)(L + 1) * n, v1); /* destroys u[1] */
                for (u32 i = 0; i < L; ++i)
                    for (u32 x = 0; x < n; ++x) tgt[(size_t)i * n + x] = v1[(size_t)i * n + tab[x]];
                ks_decompose(c, L, tgt, ext);
                ks_inner(c, L, ext, NULL, giant_keys[g], acc);
                add1 = acc + (size_t)(L + 1) * n;
            }
            for (u32 I = 0; I <= L; ++I) {
                const u64 m = c->q[I == L ? K - 1 : I];
                u64 *s0 = F + (size_t)I * n, *s1 = F + ((size_t)(L + 1) + I) * n;
                const u64 *u0 = u + (size_t)I * n, *a1 = add1 + (size_t)I * n;
                for (u32 x = 0; x < n; ++x) {
                    s0[x] = addmod(s0[x], u0[tab ? tab[x] : x], m); /* pi_g(u_g[0]) stays in the extended basis */
                    if (tab) s0[x] = addmod(s0[x], acc[(size_t)I * n + x], m);
                    s1[x] = addmod(s1[x], a1[x], m);
                }
            }
        }
        ks_mod_down_add(c, L, F, zero, res);
        if (rescale) orc_rescale(c, L, res, 2, out + (size_t)b * 2 * (L - 1) * n);
        else memcpy(out + (size_t)b * ctw, res, sizeof(u64) * ctw);
        free(ext); free(baby); free(u); free(acc); free(F); free(v1); free(zero); free(res); free(tgt); free(lazy);
    }
    free(tabs);
}
