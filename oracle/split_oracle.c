/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C) of the reference's train/test splitters,
 *   split_data_selected_users   /root/reference/src/recometrics.hpp:1015-1106
 *   split_data_separate_users   /root/reference/src/recometrics.hpp:1201-1322
 *   split_data_joined_users     /root/reference/src/recometrics.hpp:1439-1505 (+ concat_csr_matrices :1324-1359)
 * together with the pieces of libstdc++ 13 (the library the reference is compiled against in this image) that decide
 * their output: std::mt19937, uniform_int_distribution<unsigned long> over a 32-bit generator (Lemire's nearly
 * divisionless method, bits/uniform_int_dist.h `_S_nd`) and std::shuffle (bits/stl_algo.h: two swap positions per
 * draw while n*n fits the generator's range, one per draw above that).
 *
 * Parity status: PINNED.  tests/test_split_oracle.py checks it entry for entry against
 * oracle/_ref/librecometrics_ref.so (the unmodified reference, compiled by oracle/Makefile) on seeded inputs and against
 * the committed fixtures tests/golden/split_*.npz made from that library (tests/golden/make_golden_split.py).
 *
 * One documented difference: rows whose item ids repeat.  The reference orders each half of a split row with an
 * unstable std::sort on the item id, so the order of the VALUES of a repeated id is libstdc++'s; here it is the input order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- std::mt19937 (seeded with `value mod 2^32`) ---- */
typedef struct { uint32_t x[624]; int pos; } sp_mt_t;

static void sp_mt_seed(sp_mt_t *g, uint64_t value)
{
    g->x[0] = (uint32_t)value;
    for (int i = 1; i < 624; i++)
        g->x[i] = 1812433253u * (g->x[i - 1] ^ (g->x[i - 1] >> 30)) + (uint32_t)i;
    g->pos = 624;
}

static uint32_t sp_mt_next(sp_mt_t *g)
{
    if (g->pos >= 624) {
        for (int i = 0; i < 624; i++) {
            const uint32_t y = (g->x[i] & 0x80000000u) | (g->x[(i + 1) % 624] & 0x7fffffffu);
            g->x[i] = g->x[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->pos = 0;
    }
    uint32_t y = g->x[g->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* uniform integer in [0, range) for 1 <= range < 2^32: `_S_nd<uint64_t>` */
static uint32_t sp_below(sp_mt_t *g, uint32_t range)
{
    uint64_t product = (uint64_t)sp_mt_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        const uint32_t threshold = (uint32_t)(0u - range) % range;
        while (low < threshold) {
            product = (uint64_t)sp_mt_next(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (uint32_t)(product >> 32);
}

static void sp_swap(int32_t *a, int32_t *b) { const int32_t t = *a; *a = *b; *b = t; }

/* std::shuffle(v, v + len, mt19937) */
static void sp_shuffle(int32_t *v, int64_t len, sp_mt_t *g)
{
    if (len <= 0) return;
    const uint64_t urng = 0xffffffffull, n = (uint64_t)len;
    if (urng / n >= n) {
        int64_t i = 1;
        if ((n % 2) == 0) {
            sp_swap(&v[i], &v[sp_below(g, 2)]);
            i++;
        }
        while (i != len) {
            const uint64_t r = (uint64_t)i + 1;                  /* positions open to element i */
            const uint32_t x = sp_below(g, (uint32_t)(r * (r + 1)));
            sp_swap(&v[i], &v[x / (r + 1)]);
            sp_swap(&v[i + 1], &v[x % (r + 1)]);
            i += 2;
        }
        return;
    }
    for (int64_t i = 1; i < len; i++)
        sp_swap(&v[i], &v[sp_below(g, (uint32_t)(i + 1))]);   /* (a row beyond 2^32 entries cannot exist: int32 CSR) */
}

static const int32_t *g_sort_items;
static int sp_by_item(const void *a, const void *b)
{
    const int32_t ia = *(const int32_t *)a, ib = *(const int32_t *)b;
    const int32_t ka = g_sort_items[ia], kb = g_sort_items[ib];
    if (ka != kb) return ka < kb ? -1 : 1;
    return ia < ib ? -1 : (ia > ib);
}

static int sp_int_cmp(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return x < y ? -1 : (x > y);
}

/* values are moved as opaque elements of `vsz` bytes (4 or 8) */
static void sp_copy_vals(char *dst, int64_t d, const char *src, int64_t s, int64_t count, int vsz)
{
    memcpy(dst + d * vsz, src + s * vsz, (size_t)(count * vsz));
}

/* split_data_selected_users.  Outputs sized by the caller: pointers m+1, indices / values nnz.  Returns 0, or 1 for the
 * reference's "negative dimensions" error.  Not thread safe (qsort comparator state). */
int rmo_split_selected_users(const int32_t *Xp, const int32_t *Xi, const void *Xv, int vsz, int32_t m, int32_t n,
                             double test_fraction, uint64_t seed,
                             int32_t *trp, int32_t *tri, void *trv, int32_t *tep, int32_t *tei, void *tev)
{
    if (!m) return 0;
    if (m < 0 || n < 0) return 1;
    trp[0] = 0; tep[0] = 0;
    for (int32_t u = 0; u < m; u++)
        tep[u + 1] = tep[u] + (int32_t)round((double)(Xp[u + 1] - Xp[u]) * test_fraction);
    for (int32_t u = 0; u < m; u++)
        trp[u + 1] = Xp[u + 1] - tep[u + 1];

    sp_mt_t g;
    sp_mt_seed(&g, seed);
    int32_t longest = 1;
    for (int32_t u = 0; u < m; u++) if (Xp[u + 1] - Xp[u] > longest) longest = Xp[u + 1] - Xp[u];
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)longest);
    if (!idx) return 2;
    for (int32_t u = 0; u < m; u++) {
        const int32_t cnt = Xp[u + 1] - Xp[u];
        if (!cnt) continue;
        const int32_t held = tep[u + 1] - tep[u];
        const int32_t *it = Xi + Xp[u];
        if (!held) {
            memcpy(tri + trp[u], it, sizeof(int32_t) * (size_t)cnt);
            sp_copy_vals((char *)trv, trp[u], (const char *)Xv, Xp[u], cnt, vsz);
            continue;
        }
        if (held == cnt) {
            memcpy(tei + tep[u], it, sizeof(int32_t) * (size_t)cnt);
            sp_copy_vals((char *)tev, tep[u], (const char *)Xv, Xp[u], cnt, vsz);
            continue;
        }
        for (int32_t j = 0; j < cnt; j++) idx[j] = j;
        sp_shuffle(idx, cnt, &g);
        g_sort_items = it;
        qsort(idx, (size_t)held, sizeof(int32_t), sp_by_item);
        qsort(idx + held, (size_t)(cnt - held), sizeof(int32_t), sp_by_item);
        for (int32_t j = 0; j < held; j++) {
            tei[tep[u] + j] = it[idx[j]];
            sp_copy_vals((char *)tev, tep[u] + j, (const char *)Xv, Xp[u] + idx[j], 1, vsz);
        }
        for (int32_t j = held; j < cnt; j++) {
            tri[trp[u] + j - held] = it[idx[j]];
            sp_copy_vals((char *)trv, trp[u] + j - held, (const char *)Xv, Xp[u] + idx[j], 1, vsz);
        }
    }
    free(idx);
    return 0;
}

/* split_data_separate_users (joined = 0) / split_data_joined_users (joined = 1).
 * Outputs sized by the caller: users_test m; every pointer array m+1; every index / value array nnz.
 * out_sizes[0] = test users taken, [1] = rows of the remainder (separate) -- the row counts of the other outputs follow.
 * joined: the training output holds the taken users' training rows followed by the remainder; rem* are not written.
 * Returns 0, or the reference's errors: 1 "Target number of test users is larger than available users", 2 "Selected
 * minimum number of items is larger than total number of items", 3 "No users satisfy criteria for test inclusion". */
int rmo_split_users(const int32_t *Xp, const int32_t *Xi, const void *Xv, int vsz, int32_t m, int32_t n,
                    int32_t n_users_test, double test_fraction, int consider_cold_start, int32_t min_items_pool,
                    int32_t min_pos_test, uint64_t seed, int joined,
                    int32_t *users_test, int32_t *rep, int32_t *rei, void *rev,
                    int32_t *trp, int32_t *tri, void *trv, int32_t *tep, int32_t *tei, void *tev, int64_t *out_sizes)
{
    if (n_users_test > m) return 1;
    if (min_items_pool >= n) return 2;
    sp_mt_t g;
    sp_mt_seed(&g, seed);
    int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    for (int32_t u = 0; u < m; u++) ids[u] = u;
    sp_shuffle(ids, m, &g);

    int32_t end = m, taken = 0;
    do {
        const int32_t u = ids[taken];
        const int32_t cnt = Xp[u + 1] - Xp[u];
        const int32_t held = (int32_t)round((double)cnt * test_fraction);
        const int eligible = cnt != 0 && held >= min_pos_test && n - (cnt - held) >= min_items_pool &&
                             (consider_cold_start || held != cnt) && cnt + 1 < n;
        if (eligible) taken++;
        else sp_swap(&ids[taken], &ids[--end]);
    } while (taken < n_users_test && taken < end);
    if (!taken) { free(ids); return 3; }

    qsort(ids, (size_t)taken, sizeof(int32_t), sp_int_cmp);
    memcpy(users_test, ids, sizeof(int32_t) * (size_t)taken);
    qsort(ids + taken, (size_t)(m - taken), sizeof(int32_t), sp_int_cmp);

    /* the taken users' rows as a matrix of their own, split with a FRESH generator on the same seed */
    int32_t *sp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(taken + 1));
    sp[0] = 0;
    for (int32_t r = 0; r < taken; r++) sp[r + 1] = sp[r] + (Xp[ids[r] + 1] - Xp[ids[r]]);
    const int32_t sel_nnz = sp[taken];
    int32_t *si = (int32_t *)malloc(sizeof(int32_t) * (size_t)(sel_nnz > 0 ? sel_nnz : 1));
    char *sv = (char *)malloc((size_t)vsz * (size_t)(sel_nnz > 0 ? sel_nnz : 1));
    for (int32_t r = 0; r < taken; r++) {
        const int32_t cnt = sp[r + 1] - sp[r];
        memcpy(si + sp[r], Xi + Xp[ids[r]], sizeof(int32_t) * (size_t)cnt);
        sp_copy_vals(sv, sp[r], (const char *)Xv, Xp[ids[r]], cnt, vsz);
    }
    const int rc = rmo_split_selected_users(sp, si, sv, vsz, taken, n, test_fraction, seed, trp, tri, trv, tep, tei, tev);
    free(sp); free(si); free(sv);
    if (rc) { free(ids); return 10 + rc; }

    /* the other users, in ascending order: their own matrix (separate) or appended below the training rows (joined) */
    const int32_t others = m - taken;
    int32_t *op = joined ? trp + taken : rep;
    int32_t *oi = joined ? tri : rei;
    char *ov = (char *)(joined ? trv : rev);
    const int32_t base = joined ? trp[taken] : 0;
    if (!joined) op[0] = 0;
    for (int32_t r = 0; r < others; r++) {
        const int32_t u = ids[taken + r], cnt = Xp[u + 1] - Xp[u];
        const int32_t at = (r == 0) ? base : op[r];
        memcpy(oi + at, Xi + Xp[u], sizeof(int32_t) * (size_t)cnt);
        sp_copy_vals(ov, at, (const char *)Xv, Xp[u], cnt, vsz);
        op[r + 1] = at + cnt;
    }
    out_sizes[0] = taken;
    out_sizes[1] = others;
    free(ids);
    return 0;
}
