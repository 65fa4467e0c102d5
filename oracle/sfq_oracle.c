/* TEST INFRASTRUCTURE ONLY - see sfq_oracle.h.  Plain-C restatement of the reference's hot
 * path, written from the behaviour of the files cited per function (paths are relative to
 * /root/reference).  Dense tables exactly like the reference (the CUDA product uses hashed /
 * sparse layouts; the two implementations share no code).
 */
#include "sfq_oracle.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ errors */
typedef struct { char *err; int failed; } errctx;
static void fail(errctx *e, const char *fmt, ...) {
    if (e->failed) return;
    e->failed = 1;
    if (!e->err) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e->err, 256, fmt, ap);
    va_end(ap);
}

/* ------------------------------------------------------------------ byte sink */
typedef struct { uint8_t *p; size_t n, cap; } obuf;
static void ob_put(obuf *o, uint8_t c) {
    if (o->n == o->cap) {
        o->cap = o->cap ? o->cap * 2 : 256;
        o->p = (uint8_t *)realloc(o->p, o->cap);
    }
    o->p[o->n++] = c;
}

/* ------------------------------------------------------------------ range coder
 * coder.hpp:21-103.  TOP = 1<<24, 64-bit low/code, 32-bit range. */
#define RC_TOP (1u << 24)
typedef struct { uint64_t low; uint32_t range; obuf out; int live; } rc_enc;
typedef struct { uint64_t low, code; uint32_t range; const uint8_t *p; size_t n, pos; } rc_dec;

static void enc_init(rc_enc *r) { r->low = 0; r->range = 0xFFFFFFFFu; r->live = 1; }   /* :34-39 */
static void enc_done(rc_enc *r) {                                                      /* :52-61 */
    for (int i = 0; i < 8; i++) { ob_put(&r->out, (uint8_t)(r->low >> 56)); r->low <<= 8; }
}
static void enc_put(rc_enc *r, uint32_t cum, uint32_t freq, uint32_t tot) {            /* :66-81 */
    r->range /= tot;
    r->low += (uint32_t)(cum * r->range);
    r->range *= freq;
    while (r->range < RC_TOP) {
        if ((r->low ^ (r->low + r->range)) & (0xffULL << 56))
            r->range = (((uint32_t)r->low) | (RC_TOP - 1)) - (uint32_t)r->low;
        ob_put(&r->out, (uint8_t)(r->low >> 56));
        r->range <<= 8;
        r->low <<= 8;
    }
}
static uint8_t dec_byte(rc_dec *r) { return r->pos < r->n ? r->p[r->pos++] : 0; }      /* filer.hpp:94-97 */
static void dec_init(rc_dec *r, const uint8_t *p, size_t n) {                          /* :41-49 */
    r->p = p; r->n = n; r->pos = 0; r->low = 0; r->range = 0xFFFFFFFFu; r->code = 0;
    for (int i = 0; i < 8; i++) r->code = (r->code << 8) | dec_byte(r);
}
static uint32_t dec_freq(rc_dec *r, uint32_t tot) {                                    /* :83-86 */
    r->range /= tot;
    return (uint32_t)(r->code / r->range);
}
static void dec_take(rc_dec *r, uint32_t cum, uint32_t freq) {                         /* :88-102 */
    uint32_t t = cum * r->range;
    r->low += t;
    r->code -= t;
    r->range *= freq;
    while (r->range < RC_TOP) {
        if ((r->low ^ (r->low + r->range)) & (0xffULL << 56))
            r->range = (((uint32_t)r->low) | (RC_TOP - 1)) - (uint32_t)r->low;
        r->code = (r->code << 8) | dec_byte(r);
        r->range <<= 8;
        r->low <<= 8;
    }
}

/* ------------------------------------------------------------------ 4-symbol model
 * base2_ranger.hpp:35-105.  Stored XOR 0x03030303 so that calloc'ed memory is the
 * initial state (3,3,3,3). */
static inline uint32_t b2_load(const uint32_t *t) { return *t ^ 0x03030303u; }
static inline void b2_store(uint32_t *t, uint32_t v) { *t = v ^ 0x03030303u; }
static inline uint32_t b2_update(uint32_t v, int s) {                                  /* :48-53,60-66 */
    if (((v >> (8 * s)) & 0xff) > 254)
        v = ((v & ~0x01010101u) >> 1) | (v & 0x01010101u);
    return v + (1u << (8 * s));
}
static void b2_put(uint32_t *slot, rc_enc *rc, int s) {                                /* :74-84 */
    uint32_t v = b2_load(slot);
    uint32_t f[4] = { v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff, v >> 24 };
    uint32_t tot = f[0] + f[1] + f[2] + f[3], cum = 0;
    for (int j = 0; j < s; j++) cum += f[j];
    enc_put(rc, cum, f[s], tot);
    b2_store(slot, b2_update(v, s));
}
static int b2_get(uint32_t *slot, rc_dec *rc) {                                        /* :86-104 */
    uint32_t v = b2_load(slot);
    uint32_t f[4] = { v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff, v >> 24 };
    uint32_t tot = f[0] + f[1] + f[2] + f[3];
    uint32_t prob = dec_freq(rc, tot), cum = 0;
    int i;
    for (i = 0; i < 3; i++) {          /* i==4 would be an assert failure in the reference */
        if (cum + f[i] <= prob) cum += f[i]; else break;
    }
    dec_take(rc, cum, f[i]);
    b2_store(slot, b2_update(v, i));
    return i;
}

/* ------------------------------------------------------------------ 64 / 256-symbol models
 * log64_ranger.hpp:36-140 and power_ranger.hpp:36-131 are the same scheme with different
 * constants; one restatement parameterised by (nsym, step, maxfreq, slack). */
typedef struct { uint32_t total; uint16_t iend; uint8_t count; uint16_t freq[256]; uint8_t syms[256]; } amodel;
typedef struct { uint32_t total; uint16_t iend; uint8_t count; uint16_t freq[64]; uint8_t syms[64]; } amodel64;
typedef struct { int nsym, step, maxfreq, slack; } aparm;
static const aparm P_LOG64 = { 64, 6, (1 << 16) - 64, 20 };      /* log64_ranger.hpp:37-42,72 */
static const aparm P_POWER = { 256, 14, (1 << 15) - 32, 256 };   /* power_ranger.hpp:37-41,71 */

static uint8_t am_update(uint32_t *total, uint16_t *iend, uint8_t *count, uint16_t *freq,
                         uint8_t *syms, const aparm *P, uint32_t i) {
    if (freq[i] > (uint32_t)(P->maxfreq - P->step)) {            /* log64:71-77, power:70-76 */
        if (i == 0 && (uint32_t)freq[0] + (uint32_t)P->slack > *total) return syms[0];
        uint32_t t = 0;
        for (uint32_t j = 0; j < *iend; j++) t += (freq[j] >>= 1);
        *total = t;
    }
    freq[i] = (uint16_t)(freq[i] + P->step);
    *total += P->step;
    if (i == 0) return syms[0];                                   /* ++count not evaluated */
    *count = (uint8_t)(*count + 1);
    if ((*count & 0xf) || freq[i] <= freq[i - 1]) return syms[i];
    uint8_t c = syms[i]; syms[i] = syms[i - 1]; syms[i - 1] = c;  /* down_level */
    uint16_t f = freq[i]; freq[i] = freq[i - 1]; freq[i - 1] = f;
    return c;
}
static void am_put(uint32_t *total, uint16_t *iend, uint8_t *count, uint16_t *freq, uint8_t *syms,
                   const aparm *P, rc_enc *rc, uint32_t sym) {   /* log64:98-112, power:93-106 */
    while (*iend <= sym) { syms[*iend] = (uint8_t)*iend; (*iend)++; }
    uint32_t i = 0, sumf = 0;
    for (; syms[i] != sym; i++) sumf += freq[i];
    enc_put(rc, sumf + i, (uint32_t)freq[i] + 1, *total + P->nsym);
    am_update(total, iend, count, freq, syms, P, i);
}
static uint32_t am_get(uint32_t *total, uint16_t *iend, uint8_t *count, uint16_t *freq, uint8_t *syms,
                       const aparm *P, rc_dec *rc) {             /* log64:114-138, power:108-130 */
    uint32_t vtot = *total + P->nsym, sumf = 0, i;
    uint32_t prob = dec_freq(rc, vtot);
    for (i = 0; i < (uint32_t)P->nsym; i++) {
        if (*iend == i) { syms[*iend] = (uint8_t)i; (*iend)++; }
        if (sumf + freq[i] + 1 <= prob) sumf += freq[i] + 1; else break;
    }
    if (i >= (uint32_t)P->nsym) i = P->nsym - 1;                 /* corrupt stream; stay in bounds */
    dec_take(rc, sumf, (uint32_t)freq[i] + 1);
    return am_update(total, iend, count, freq, syms, P, i);
}
#define PW_PUT(m, rc, s) am_put(&(m)->total, &(m)->iend, &(m)->count, (m)->freq, (m)->syms, &P_POWER, rc, s)
#define PW_GET(m, rc)    am_get(&(m)->total, &(m)->iend, &(m)->count, (m)->freq, (m)->syms, &P_POWER, rc)
#define L64_PUT(m, rc, s) am_put(&(m)->total, &(m)->iend, &(m)->count, (m)->freq, (m)->syms, &P_LOG64, rc, s)
#define L64_GET(m, rc)    am_get(&(m)->total, &(m)->iend, &(m)->count, (m)->freq, (m)->syms, &P_LOG64, rc)

/* PowerRangerU, power_ranger.hpp:133-192 */
typedef struct { amodel p[14]; } umodel;
static void pu_put(umodel *u, rc_enc *rc, uint64_t num) {
    if (num <= 0x7f) { PW_PUT(&u->p[0], rc, (uint32_t)(num & 0xff)); return; }
    if (num < 0x7ffe) {
        PW_PUT(&u->p[0], rc, (uint32_t)(0xff & (0x80 | (num >> 8))));
        PW_PUT(&u->p[1], rc, (uint32_t)(0xff & num));
        return;
    }
    PW_PUT(&u->p[0], rc, 0xff);
    if (num < (1ULL << 32)) {
        PW_PUT(&u->p[1], rc, 0xfe);
        for (int sh = 0, i = 2; sh < 32; sh += 8, i++) PW_PUT(&u->p[i], rc, (uint32_t)(0xff & (num >> sh)));
    } else {
        PW_PUT(&u->p[1], rc, 0xff);
        for (int sh = 0, i = 6; sh < 64; sh += 8, i++) PW_PUT(&u->p[i], rc, (uint32_t)(0xff & (num >> sh)));
    }
}
static uint64_t pu_get(umodel *u, rc_dec *rc) {
    uint64_t num = PW_GET(&u->p[0], rc);
    if (num > 0x7f) {
        num = (num << 8) | PW_GET(&u->p[1], rc);
        if (num < 0xfffe) num &= 0x7fff;
        else if (num == 0xfffe) {
            num = 0;
            for (int sh = 0, i = 2; sh < 32; sh += 8, i++) num |= (uint64_t)PW_GET(&u->p[i], rc) << sh;
        } else {
            num = 0;
            for (int sh = 0, i = 6; sh < 64; sh += 8, i++) num |= (uint64_t)PW_GET(&u->p[i], rc) << sh;
        }
    }
    return num;
}

/* ------------------------------------------------------------------ exception-list streams
 * xfile.cpp:36-110: created on first put; closing emits put(0) then the 8-byte flush. */
typedef struct { rc_enc rc; umodel num; amodel str; } xsave;
typedef struct { rc_dec rc; umodel num; amodel str; int valid; } xload;
static void xs_put(xsave *x, uint64_t v) { if (!x->rc.live) enc_init(&x->rc); pu_put(&x->num, &x->rc, v); }
static void xs_put_chr(xsave *x, uint8_t c) { if (!x->rc.live) enc_init(&x->rc); PW_PUT(&x->str, &x->rc, c); }
static void xs_put_str(xsave *x, const uint8_t *p, size_t len) {
    xs_put(x, len);
    for (size_t j = 0; j < len; j++) PW_PUT(&x->str, &x->rc, p[j]);
}
static void xs_close(xsave *x) { if (x->rc.live) { xs_put(x, 0); enc_done(&x->rc); } }
static void xl_open(xload *x, const uint8_t *p, size_t n) {
    memset(x, 0, sizeof *x);
    x->valid = (p != NULL && n > 0);
    if (x->valid) dec_init(&x->rc, p, n);
}
static uint64_t xl_get(xload *x) { return x->valid ? pu_get(&x->num, &x->rc) : 0; }
static uint8_t xl_get_chr(xload *x) { return (uint8_t)PW_GET(&x->str, &x->rc); }

/* ------------------------------------------------------------------ quality model
 * qlts.hpp:35-74, qlts.cpp:34-39, 74-136 (save), 163-234 (load). */
typedef struct { amodel64 *ranger; amodel ex; uint32_t mask; int level; uint32_t extra_hi; } qmodel;
static int q_init(qmodel *q, int level) {
    memset(q, 0, sizeof *q);
    q->level = level;
    size_t cnt = level == 1 ? (1u << 12) : (1u << 16);
    q->mask = (uint32_t)cnt - 1;
    q->ranger = (amodel64 *)calloc(cnt, sizeof(amodel64));
    return q->ranger != NULL;
}
typedef struct { uint32_t last, delta, di; uint8_t q1, q2; } qctx;
static void qctx_reset(qctx *c) { c->last = 0; c->delta = 5; c->di = 0; c->q1 = c->q2 = 0; }
static uint32_t q_delta_ctx(uint32_t *delta, uint8_t q, uint8_t q1, uint8_t q2) {      /* qlts.hpp:62-74 */
    if (q1 > q) *delta += (uint32_t)(q1 - q);
    uint32_t d = *delta >> 3;
    return ((uint32_t)q | ((uint32_t)(q1 < q2 ? q2 : q1) << 6) | ((uint32_t)(q1 == q2) << 12)
            | ((7 > d ? d : 7) << 13)) & 0xFFFF;
}
static void qctx_next(const qmodel *q, qctx *c, uint8_t b) {
    if (q->level <= 2) { c->last = ((uint32_t)b | (c->last << 6)) & q->mask; return; }  /* hpp:52-57 */
    if (++c->di & 1) { c->last = q_delta_ctx(&c->delta, b, c->q1, c->q2); c->q2 = b; } /* cpp:127-134 */
    else             { c->last = q_delta_ctx(&c->delta, b, c->q2, c->q1); c->q1 = b; }
}
static void q_save(qmodel *q, rc_enc *rc, const uint8_t *buf, size_t size) {
    qctx c; qctx_reset(&c);
    for (size_t k = 0; k < size; k++) {
        uint8_t b = (uint8_t)(buf[k] - '!');
        amodel64 *m = &q->ranger[c.last];
        if (b < 63) L64_PUT(m, rc, b);
        else { L64_PUT(m, rc, 63); PW_PUT(&q->ex, rc, b); q->extra_hi++; }
        qctx_next(q, &c, b);
    }
}
static void q_load(qmodel *q, rc_dec *rc, uint8_t *buf, size_t size) {
    qctx c; qctx_reset(&c);
    for (size_t k = 0; k < size; k++) {
        amodel64 *m = &q->ranger[c.last];
        uint8_t b = (uint8_t)L64_GET(m, rc);
        if (b == 63) b = (uint8_t)PW_GET(&q->ex, rc);
        buf[k] = (uint8_t)('!' + b);
        qctx_next(q, &c, b);
    }
    buf[size] = '\n';
}

/* ------------------------------------------------------------------ base model
 * gens.hpp:43-82, gens.cpp:36-44, 72-77, 91-159 (save), 164-249 (load). */
typedef struct {
    uint32_t *ranger; uint32_t mask;
    uint64_t genofs, ns_index, nn_index;
    uint8_t n_byte;
} gmodel;
static int g_init(gmodel *g, int level) {
    memset(g, 0, sizeof *g);
    int bits = level == 1 ? 18 : level == 2 ? 22 : level == 3 ? 24 : 26;
    g->mask = (1u << bits) - 1;
    g->ranger = (uint32_t *)calloc((size_t)1 << bits, 4);
    return g->ranger != NULL;
}
static int gencode(uint8_t c) {                                                        /* gens.cpp:72-77 */
    switch (c) {
    case '0': case 'A': case 'a': return 0;
    case '1': case 'C': case 'c': return 1;
    case '2': case 'G': case 'g': return 2;
    case '3': case 'T': case 't': return 3;
    case '.': case 'N': case 'n': return 4;
    default: return 0x10;
    }
}
static void g_save(gmodel *g, rc_enc *rc, xsave *x_ns, xsave *x_nn, const uint8_t *gen,
                   const uint8_t *qlt, uint64_t llen, uint64_t qlen, errctx *e) {
    uint32_t last = 0x007616c7;                                                        /* :139 */
    for (uint64_t i = 0; i < llen && !e->failed; i++) {
        uint8_t qc = i < qlen ? qlt[i] : 40;                                           /* :153 */
        int n = gencode(gen[i]);
        int bad_n = 0, bad_q = (qc == '!');
        if (n > 3) {
            if (n > 4) { fail(e, "unexpected genome char: %c", gen[i]); return; }
            bad_n = 1; n = 0;
        }
        g->genofs++;
        if (bad_n || bad_q) {                                                          /* :91-114 */
            if (!bad_n) { xs_put(x_nn, g->genofs - g->nn_index); g->nn_index = g->genofs; }
            else {
                if (!g->n_byte) g->n_byte = gen[i];
                if (gen[i] != g->n_byte) { fail(e, "switched N_byte: %c", gen[i]); return; }
                if (!bad_q) { xs_put(x_ns, g->genofs - g->ns_index); g->ns_index = g->genofs; }
            }
        }
        last &= g->mask;
        b2_put(&g->ranger[last], rc, n);
        last = (last << 2) | (uint32_t)n;
    }
}
static void g_load(gmodel *g, rc_dec *rc, xload *x_ns, xload *x_nn, const char *alphabet,
                   uint8_t *gen, const uint8_t *qlt, uint64_t llen, uint64_t qlen) {
    uint32_t last = 0x007616c7;
    for (uint64_t i = 0; i < llen; i++) {
        last &= g->mask;
        int b = b2_get(&g->ranger[last], rc);
        gen[i] = (uint8_t)alphabet[b];
        last = (last << 2) + (uint32_t)b;
        uint8_t qc = i < qlen ? qlt[i] : 40;
        g->genofs++;                                                                   /* :200-213 */
        if (g->nn_index == g->genofs) g->nn_index += xl_get(x_nn);
        else if (qc == '!') gen[i] = g->n_byte;
        else if (g->ns_index == g->genofs) { gen[i] = g->n_byte; g->ns_index += xl_get(x_ns); }
    }
}

/* ------------------------------------------------------------------ header model
 * recs.hpp:42-79, recs.cpp:139-262 (tokeniser, numberwang), 277-372 (save), 374-461 (load). */
typedef struct { amodel type, str; umodel num; } hranger;
typedef struct { int off[66]; int wln[66]; uint8_t str[66]; int len; } spacemap;
typedef struct {
    hranger *ranger;           /* [66] */
    spacemap smap[2];
    uint8_t ctype[2][65];
    uint64_t cnumb[2][65];
    int imap, initialized;
    uint64_t index;            /* m_last.index */
    int pre5;                  /* header stream of format versions < 5 (recs.cpp:397-398, 463-510) */
} hmodel;
enum { ST_DGT = 0, ST_DLT, ST_STR, ST_HGT, ST_HLT, ST_HGT_Z, ST_HLT_Z, ST_HGTC, ST_HLTC,
       ST_HGTC_Z, ST_HLTC_Z, ST_DGT_Z, ST_DLT_Z };
static int is_dig(uint8_t c) { return c >= '0' && c <= '9'; }
static int is_word(uint8_t c) { return is_dig(c) || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }

static int h_init(hmodel *h) {
    memset(h, 0, sizeof *h);
    h->ranger = (hranger *)calloc(66, sizeof(hranger));
    return h->ranger != NULL;
}
static void map_space(spacemap *m, const uint8_t *p, errctx *e) {                      /* :140-157 */
    m->len = 0;
    m->off[0] = 0;
    for (int i = 0;; i++) {
        if (!is_word(p[i])) {
            m->wln[m->len] = i - m->off[m->len];
            m->str[m->len++] = p[i];
            m->off[m->len] = i + 1;
            if (p[i] == 0 || p[i] == '\n') break;
            if (m->len > 64) { fail(e, "ERROR: irregulal record (over 64 non alpha non digit). Is it a valid fastq file?"); return; }
        }
    }
}
static int numberwang(const uint8_t *p, int len, uint64_t *num, uint8_t pctype) {      /* :192-262 */
    int i = 0;
    int has_z = p[i] == '0';
    if (has_z && p[++i] == '0') return ST_STR;
    int caps = 0;
    *num = 0;
    while (pctype != 2) {
        if (i >= len) return has_z ? ST_DGT_Z : ST_DGT;
        if (is_dig(p[i])) {
            uint64_t t = (*num << 3) + (*num << 1) + p[i++] - '0';
            if (t < *num) return ST_STR;
            *num = t;
            continue;
        }
        if ((p[i] | 0x20) < 'a' || (p[i] | 0x20) > 'f') return ST_STR;
        caps = 1 + (p[i] < 'a');
        i = has_z;
        *num = 0;
        break;
    }
    if (len > 16) return ST_STR;
    for (; i < len; i++) {
        int nib;
        if (is_dig(p[i])) nib = p[i] - '0';
        else if (p[i] >= 'a' && p[i] <= 'f') { if (caps == 2) return ST_STR; caps = 1; nib = 10 + p[i] - 'a'; }
        else if (p[i] >= 'A' && p[i] <= 'F') { if (caps == 1) return ST_STR; caps = 2; nib = 10 + p[i] - 'A'; }
        else return ST_STR;
        *num = (*num << 4) + (uint64_t)nib;
    }
    return caps == 2 ? (has_z ? ST_HGTC_Z : ST_HGTC) : (has_z ? ST_HGT_Z : ST_HGT);
}
static int is_number(const uint8_t *p, int len, long long *num) {                      /* recs.cpp:265-275 */
    if (*p == '0') return 0;
    *num = 0;
    for (int i = 0; i < len; i++) {
        if (!is_dig(p[i])) return 0;
        *num = (long long)(((unsigned long long)*num << 3) + ((unsigned long long)*num << 1) + (unsigned long long)(p[i] - '0'));
    }
    return 1;
}
/* the fields of a changed header as load_pre5 expects them (test-side encoder, see sfq_oracle.h) */
static void h_save_pre5_fields(hmodel *h, rc_enc *rc, const spacemap *S, const spacemap *Pm, uint64_t map,
                               const uint8_t *buf, const uint8_t *prev) {
    for (int i = 0; i < S->len; i++) {
        if (!(map & (1ULL << (i & 63)))) continue;
        const uint8_t *b = buf + S->off[i];
        hranger *R = &h->ranger[i + 1];
        long long pval = 0, cval = 0;
        int numeric = is_number(prev + Pm->off[i], Pm->wln[i], &pval) && Pm->wln[i] < 18 && S->wln[i] > 0 && S->wln[i] < 18;
        if (numeric) {                     /* what "%lld" prints must be the field itself: digits, no leading zero except "0" */
            if (b[0] == '0') numeric = S->wln[i] == 1;
            for (int j = 0; j < S->wln[i] && numeric; j++) { if (!is_dig(b[j])) numeric = 0; else cval = cval * 10 + (b[j] - '0'); }
        }
        if (!numeric) {
            PW_PUT(&R->type, rc, (uint32_t)ST_STR);
            pu_put(&R->num, rc, (uint64_t)S->wln[i]);
            for (int j = 0; j < S->wln[i]; j++) PW_PUT(&R->str, rc, b[j]);
            continue;
        }
        PW_PUT(&R->type, rc, (uint32_t)(cval >= pval ? ST_DGT : ST_DLT));
        pu_put(&R->num, rc, (uint64_t)(cval >= pval ? cval - pval : pval - cval));
    }
}
static void h_save(hmodel *h, rc_enc *rc, xsave *x_rec, uint64_t recno, const uint8_t *buf,
                   const uint8_t *end, const uint8_t *prev, sfq_or_chunk *out, errctx *e) {
    if (!h->initialized) {                                                             /* :279-286, 68-75 */
        size_t n = (size_t)(end - buf), k;
        if (n > 399) { fail(e, "first header longer than 399 chars overflows the reference's buffer (recs.cpp:31,69)"); return; }
        for (k = 0; k < n && buf[k]; k++) out->rec_first[k] = (char)buf[k];
        out->rec_first[k] = 0;
        out->rec_first_len = (uint32_t)k;
        h->initialized = 1;
        h->imap = 0;
        map_space(&h->smap[0], buf, e);
        memset(h->ctype, 0, sizeof h->ctype);
        return;
    }
    int pm = h->imap, im = h->imap = !h->imap;
    spacemap *S = &h->smap[im], *Pm = &h->smap[pm];
    map_space(S, buf, e);
    if (e->failed) return;
    if (S->len != Pm->len || memcmp(S->str, Pm->str, (size_t)S->len)) {               /* :291-304 */
        xs_put(x_rec, recno - h->index);
        h->index = recno;
        xs_put_str(x_rec, buf, (size_t)(end - buf));
        memset(h->ctype[im], 0, sizeof h->ctype[im]);
        return;
    }
    uint64_t map = 0;
    for (int i = 0; i < S->len; i++)
        if (S->wln[i] != Pm->wln[i] || memcmp(buf + S->off[i], prev + Pm->off[i], (size_t)S->wln[i]))
            map |= 1ULL << (i & 63);                    /* x86 shift semantics of DO_SET at i==64 */
    pu_put(&h->ranger[0].num, rc, map);                                               /* :312 */
    if (h->pre5) { h_save_pre5_fields(h, rc, S, Pm, map, buf, prev); return; }
    for (int i = 0; i < S->len; i++) {
        if (!(map & (1ULL << (i & 63)))) {
            h->ctype[im][i] = h->ctype[pm][i];
            h->cnumb[im][i] = h->cnumb[pm][i];
            continue;
        }
        const uint8_t *b = buf + S->off[i];
        uint64_t bnum;
        int type = numberwang(b, S->wln[i], &bnum, h->ctype[pm][i]);
        hranger *R = &h->ranger[i + 1];
        if (type == ST_STR) {
            PW_PUT(&R->type, rc, (uint32_t)type);
            pu_put(&R->num, rc, (uint64_t)S->wln[i]);
            for (int j = 0; j < S->wln[i]; j++) PW_PUT(&R->str, rc, b[j]);
            h->ctype[im][i] = 0;
            continue;
        }
        uint64_t pnum = h->ctype[pm][i] ? h->cnumb[pm][i] : 0, gap;
        h->ctype[im][i] = (type < ST_STR || type >= ST_DGT_Z) ? 1 : 2;
        h->cnumb[im][i] = bnum;
        if (bnum < pnum) { gap = pnum - bnum; type++; } else gap = bnum - pnum;
        PW_PUT(&R->type, rc, (uint32_t)type);
        pu_put(&R->num, rc, gap);
    }
}
static size_t fmt_u64(uint8_t *b, uint64_t v, int base, int upper, int is_signed) {
    /* what sprintf("%lld" / "%llx" / "%llX") prints for a non-zero value (recs.cpp:436-456) */
    char tmp[24]; int n = 0; size_t o = 0;
    if (is_signed && (int64_t)v < 0) { b[o++] = '-'; v = (uint64_t)0 - v; }
    while (v) { int d = (int)(v % (uint64_t)base); tmp[n++] = (char)(d < 10 ? '0' + d : (upper ? 'A' : 'a') + d - 10); v /= (uint64_t)base; }
    while (n) b[o++] = (uint8_t)tmp[--n];
    return o;
}
static size_t h_load(hmodel *h, rc_dec *rc, xload *x_rec, uint64_t recno, uint8_t *buf,
                     const uint8_t *prev, const sfq_or_chunk *in, errctx *e) {
    if (!h->initialized) {                                                             /* :375-381 */
        h->initialized = 1;
        memset(h->ctype, 0, sizeof h->ctype);
        h->imap = 0;
        memcpy(buf, in->rec_first, in->rec_first_len);
        return in->rec_first_len;
    }
    int pm = h->imap, im = h->imap = !h->imap;
    if (h->index == recno) {                                                           /* :386-393 */
        size_t len = (size_t)xl_get(x_rec);
        for (size_t j = 0; j < len && j < 0x2000; j++) buf[j] = xl_get_chr(x_rec);
        h->index += xl_get(x_rec);
        memset(h->ctype[im], 0, sizeof h->ctype[im]);
        return len;
    }
    spacemap *S = &h->smap[0];
    map_space(S, prev, e);
    if (e->failed) return 0;
    uint64_t map = pu_get(&h->ranger[0].num, rc);
    uint8_t *b = buf;
    if (h->pre5) {                                                                     /* load_pre5, :463-510 */
        for (int i = 0; i < S->len; i++) {
            if (map & (1ULL << (i & 63))) {
                hranger *R = &h->ranger[i + 1];
                int type = (int)PW_GET(&R->type, rc);
                if (type == ST_DGT || type == ST_DLT) {
                    long long pval = 0;
                    if (!is_number(prev + S->off[i], S->wln[i], &pval)) { fail(e, "REC: pre-v5 numeric field after a non-number"); return 0; }   /* assert(expect_num) */
                    long long gap = (long long)pu_get(&R->num, rc);
                    long long val = type == ST_DGT ? (long long)((unsigned long long)pval + (unsigned long long)gap) : (long long)((unsigned long long)pval - (unsigned long long)gap);
                    if (val == 0) *b++ = '0'; else b += fmt_u64(b, (uint64_t)val, 10, 0, 1);
                } else if (type == ST_STR) {
                    uint32_t len = (uint32_t)pu_get(&R->num, rc);
                    for (uint32_t j = 0; j < len && (size_t)(b - buf) < 0x1f00; j++) *b++ = (uint8_t)PW_GET(&R->str, rc);
                } else { fail(e, "REC: bad type value %d", type); return 0; }
            } else {
                memcpy(b, prev + S->off[i], (size_t)S->wln[i]);
                b += S->wln[i];
            }
            *b++ = S->str[i];
            if ((size_t)(b - buf) > 0x1f00) { fail(e, "decoded header too long"); return 0; }
        }
        return (size_t)(b - buf) - 1;
    }
    for (int i = 0; i < S->len; i++) {
        if (!(map & (1ULL << (i & 63)))) {
            memcpy(b, prev + S->off[i], (size_t)S->wln[i]);
            b += S->wln[i];
            *b++ = S->str[i];
            h->ctype[im][i] = h->ctype[pm][i];
            h->cnumb[im][i] = h->cnumb[pm][i];
            continue;
        }
        hranger *R = &h->ranger[i + 1];
        int type = (int)PW_GET(&R->type, rc);
        if (type == ST_STR) {
            uint32_t len = (uint32_t)pu_get(&R->num, rc);
            for (uint32_t j = 0; j < len && (size_t)(b - buf) < 0x1f00; j++) *b++ = (uint8_t)PW_GET(&R->str, rc);
            h->ctype[im][i] = 0;
            *b++ = S->str[i];
            continue;
        }
        if (type > ST_DLT_Z) { fail(e, "REC: bad type value %d", type); return 0; }
        uint64_t pval = h->ctype[pm][i] == 0 ? 0 : h->cnumb[pm][i];
        uint64_t gap = pu_get(&R->num, rc);
        int less = (type == ST_DLT || type == ST_HLT || type == ST_HLT_Z || type == ST_HLTC ||
                    type == ST_HLTC_Z || type == ST_DLT_Z);
        uint64_t val = less ? pval - gap : pval + gap;
        h->ctype[im][i] = (type < ST_STR || type >= ST_DGT_Z) ? 1 : 2;
        h->cnumb[im][i] = val;
        if (val == 0) *b++ = '0';                                                      /* :453-454 */
        else {
            int dec = (type <= ST_DLT || type >= ST_DGT_Z);
            int zed = (type == ST_HGT_Z || type == ST_HLT_Z || type == ST_HGTC_Z || type == ST_HLTC_Z ||
                       type == ST_DGT_Z || type == ST_DLT_Z);
            int upper = (type >= ST_HGTC && type <= ST_HLTC_Z);
            if (zed) *b++ = '0';
            b += fmt_u64(b, val, dec ? 10 : 16, upper, dec);
        }
        *b++ = S->str[i];
        if ((size_t)(b - buf) > 0x1f00) { fail(e, "decoded header too long"); return 0; }
    }
    return (size_t)(b - buf) - 1;
}

/* ------------------------------------------------------------------ record framing + drivers */
static const char *const k_names[SFQ_OR_NSTREAMS] = {
    "rec", "gen", "qlt", "gen.Ns", "gen.Nn", "rec.x", "usr.x", "usr.x.q", "usr.pfg", "usr.pfq",
    "usr.lrec", "usr.lgen", "usr.lqlt"
};
#define NXS 10      /* exception-list streams: gen.Ns gen.Nn rec.x usr.x usr.x.q usr.pfg usr.pfq usr.lrec usr.lgen usr.lqlt */
const char *sfq_oracle_stream_name(int id) { return id >= 0 && id < SFQ_OR_NSTREAMS ? k_names[id] : ""; }

#define MAX_ID_LLEN 0x2000
#define MAX_GN_LLEN 0x10000

/* scan to '\n' the way usrs.cpp does (`sanity = LIMIT; while (--sanity && c != '\n')`): returns the
 * index of the newline, or -1 = ran off the buffer, -2 = line of LIMIT-1 or more chars (oversized). */
static long long scan_line(const uint8_t *buf, size_t n, size_t from, int limit) {
    size_t cur = from;
    int sanity = limit;
    while (--sanity) {
        if (cur >= n) return -1;
        if (buf[cur] == '\n') return (long long)cur;
        cur++;
    }
    return -2;
}

static int encode_impl(const uint8_t *buf, size_t n, int level, sfq_or_chunk *out, char *err, int pre5);
int sfq_oracle_encode(const uint8_t *buf, size_t n, int level, sfq_or_chunk *out, char *err) { return encode_impl(buf, n, level, out, err, 0); }
int sfq_oracle_encode_pre5(const uint8_t *buf, size_t n, int level, sfq_or_chunk *out, char *err) { return encode_impl(buf, n, level, out, err, 1); }
static int encode_impl(const uint8_t *buf, size_t n, int level, sfq_or_chunk *out, char *err, int pre5) {
    errctx e = { err, 0 };
    if (err) err[0] = 0;
    memset(out, 0, sizeof *out);
    level = level > 4 ? 4 : level < 1 ? 1 : level;                                    /* config.cpp:231-236 */
    out->level = level;

    qmodel Q; gmodel G; hmodel H;
    rc_enc rc_rec, rc_gen, rc_qlt;
    xsave *xs = (xsave *)calloc(NXS, sizeof(xsave));
    memset(&rc_rec, 0, sizeof rc_rec); memset(&rc_gen, 0, sizeof rc_gen); memset(&rc_qlt, 0, sizeof rc_qlt);
    int okq = q_init(&Q, level), okg = g_init(&G, level), okh = h_init(&H);
    if (!xs || !okq || !okg || !okh) { fail(&e, "out of memory"); goto done; }
    H.pre5 = pre5;
    out->version = pre5 ? 4 : 6;
    enc_init(&rc_rec); enc_init(&rc_gen); enc_init(&rc_qlt);
    xsave *x_ns = &xs[0], *x_nn = &xs[1], *x_rec = &xs[2], *x_llen = &xs[3], *x_qlen = &xs[4],
          *x_sgen = &xs[5], *x_sqlt = &xs[6], *x_lrec = &xs[7], *x_lgen = &xs[8], *x_lqlt = &xs[9];
    uint64_t recno = 0, i_long = 0;
    size_t cur = 0;

    /* get_oversized_record, usrs.cpp:269-301: the record starting at `at` goes, line by line and newline included,
     * to usr.lrec (id line without its '@', and the '+' line), usr.lgen and usr.lqlt */
#define OVERSIZED(at)                                                                                   \
    do {                                                                                                \
        xs_put(x_lrec, recno - i_long); i_long = recno;                                                 \
        size_t c_ = (at);                                                                               \
        if (buf[c_++] != '@') { fail(&e, "record %llu: bad (long) record", (unsigned long long)recno); break; } \
        xsave *dst_[4] = { x_lrec, x_lgen, x_lrec, x_lqlt };                                            \
        for (int l_ = 0; l_ < 4 && !e.failed; l_++) {                                                   \
            for (;;) {                                                                                  \
                if (c_ >= n) { fail(&e, "record %llu: seems truncated", (unsigned long long)recno); break; } \
                const uint8_t ch_ = buf[c_++];                                                          \
                xs_put_chr(dst_[l_], ch_);                                                              \
                if (ch_ == '\n') break;                                                                 \
            }                                                                                           \
        }                                                                                               \
        cur = c_;                                                                                       \
    } while (0)

    /* ---- determine_record, usrs.cpp:186-267 (oversized leading records are put away first) */
    int m_llen = 0, m_solid = 0;
    if (n == 0) { fail(&e, "no records were found"); goto done; }
    for (;;) {
        if (cur >= n) break;                           /* "all records were oversized": no llen / usr.* keys */
        if (buf[cur] != '@') { fail(&e, "first record: Missing prefix '@', is it really a fastq format?"); goto done; }
        long long q = scan_line(buf, n, cur + 1, MAX_ID_LLEN);
        if (q == -2) { recno++; OVERSIZED(cur); if (e.failed) goto done; continue; }
        if (q < 0) { fail(&e, "fastq file: record seems truncated  after record 0"); goto done; }
        size_t qg = (size_t)q + 1;
        for (int i = 1; i < MAX_GN_LLEN && !m_llen; i++) {
            if (qg + (size_t)i >= n) break;
            if (buf[qg + (size_t)i] == '\n') m_llen = i;
        }
        if (!m_llen) {
            if (qg + MAX_GN_LLEN > n) { fail(&e, "oversized or truncated first record"); goto done; }
            recno++; OVERSIZED(cur); if (e.failed) goto done; continue;
        }
        size_t p = qg + (size_t)m_llen + 1;
        if (p >= n || buf[p] != '+') { fail(&e, "first record: Missing 2nd prefix '+', is it really a fastq format?"); goto done; }
        int has2 = 0;
        for (;;) {
            if (++p >= n) { fail(&e, "fastq file: record seems truncated  after record 0"); goto done; }
            if (buf[p] == '\n') break;
            if (buf[p] != ' ') has2 = 1;
        }
        int d_solid = 0;
        for (int i = 1; i < m_llen && !d_solid && !m_solid; i++)
            switch (buf[qg + (size_t)i] | 0x20) {
            case '0': case '1': case '2': case '3': m_solid = 1; break;
            case 'a': case 'c': case 'g': case 't': d_solid = 1; break;
            default: break;
            }
        if (m_solid) m_llen--;
        out->solid = m_solid;
        out->llen = m_llen;
        out->two_id = has2;
        break;
    }

    /* ---- encode loop, usrs.cpp:392-407 with get_record :303-390 */
    uint64_t i_llen = 0, i_qlen = 0, i_sgen = 0, i_sqlt = 0;
    uint8_t pf_gen = 0, pf_qlt = 0;
    const uint8_t *rec = NULL, *prev_rec = NULL;
    for (;;) {
        ++recno;
    next_record:
        if (e.failed || cur >= n) break;
        size_t currec = cur;
        if (buf[cur++] != '@') { fail(&e, "fastq file: expecting '@', got '%c' after record %llu", buf[cur - 1], (unsigned long long)recno); break; }
        long long nl = scan_line(buf, n, cur, MAX_ID_LLEN);
        if (nl == -2) { OVERSIZED(currec); recno++; goto next_record; }
        if (nl < 0) { fail(&e, "fastq file: record seems truncated  after record %llu", (unsigned long long)recno); break; }
        const uint8_t *rec_end = buf + nl;
        cur = (size_t)nl + 1;
        uint8_t upd_pf = 0;
        if (m_solid) {
            if (cur >= n) { fail(&e, "fastq file: record seems truncated  after record %llu", (unsigned long long)recno); break; }
            if (pf_gen != buf[cur]) upd_pf = buf[cur];
            cur++;
        }
        size_t gi = cur;
        nl = scan_line(buf, n, cur, MAX_GN_LLEN);
        if (nl == -2) { OVERSIZED(currec); recno++; goto next_record; }
        if (nl < 0) { fail(&e, "fastq file: record seems truncated  after record %llu", (unsigned long long)recno); break; }
        cur = (size_t)nl;
        if (upd_pf) {                                                                  /* update(ET_SOLPF_GEN) :140-145 */
            xs_put(x_sgen, recno - i_sgen); xs_put_chr(x_sgen, upd_pf); i_sgen = recno; pf_gen = upd_pf;
        }
        if (m_llen != (int)(cur - gi)) {                                               /* update(ET_LLEN) :126-131 */
            xs_put(x_llen, recno - i_llen); xs_put(x_llen, (uint16_t)(cur - gi)); i_llen = recno;
            m_llen = (int)(uint16_t)(cur - gi);
        }
        cur++;
        if (cur >= n || buf[cur] != '+') { fail(&e, "fastq file: expecting '+', got '%c' after record %llu", cur < n ? buf[cur] : '?', (unsigned long long)recno); break; }
        cur++;
        nl = scan_line(buf, n, cur, MAX_ID_LLEN);
        if (nl == -2) { fail(&e, "wierd second id at record %llu", (unsigned long long)recno); break; }
        if (nl < 0) { fail(&e, "fastq file: record seems truncated  after record %llu", (unsigned long long)recno); break; }
        cur = (size_t)nl + 1;
        if (m_solid) {
            if (cur >= n) { fail(&e, "fastq file: record seems truncated  after record %llu", (unsigned long long)recno); break; }
            if (pf_qlt != buf[cur]) {                                                  /* update(ET_SOLPF_QLT) :147-152 */
                xs_put(x_sqlt, recno - i_sqlt); xs_put_chr(x_sqlt, buf[cur]); i_sqlt = recno; pf_qlt = buf[cur];
            }
            cur++;
        }
        size_t qi = cur;
        nl = scan_line(buf, n, cur, MAX_GN_LLEN);
        if (nl == -2) { OVERSIZED(currec); recno++; goto next_record; }      /* (the updates above have happened, as in the reference) */
        if (nl < 0) { fail(&e, "fastq file: record seems truncated  after record %llu", (unsigned long long)recno); break; }
        int m_qlen = (int)((size_t)nl - qi);
        if (m_qlen != m_llen) {                                                        /* update(ET_QLEN) :133-138 */
            xs_put(x_qlen, recno - i_qlen); xs_put(x_qlen, (uint16_t)m_qlen); i_qlen = recno;
        }
        cur = (size_t)nl + 1;
        prev_rec = rec;
        rec = buf + currec + 1;

        g_save(&G, &rc_gen, x_ns, x_nn, buf + gi, buf + qi, (uint64_t)m_llen, (uint64_t)m_qlen, &e);
        if (e.failed) break;
        h_save(&H, &rc_rec, x_rec, recno, rec, rec_end, prev_rec, out, &e);
        if (e.failed) break;
        q_save(&Q, &rc_qlt, buf + qi, (size_t)m_qlen);
    }
    if (e.failed) goto done;
    out->num_records = recno - 1;
    out->n_byte = (G.n_byte && G.n_byte != 'N') ? G.n_byte : 0;
    out->extra_hi_qlt = Q.extra_hi;
    enc_done(&rc_rec); enc_done(&rc_gen); enc_done(&rc_qlt);
    for (int k = 0; k < NXS; k++) xs_close(&xs[k]);
    out->data[SFQ_OR_REC] = rc_rec.out.p; out->size[SFQ_OR_REC] = rc_rec.out.n; rc_rec.out.p = NULL;
    out->data[SFQ_OR_GEN] = rc_gen.out.p; out->size[SFQ_OR_GEN] = rc_gen.out.n; rc_gen.out.p = NULL;
    out->data[SFQ_OR_QLT] = rc_qlt.out.p; out->size[SFQ_OR_QLT] = rc_qlt.out.n; rc_qlt.out.p = NULL;
    {
        static const int sid[NXS] = { SFQ_OR_GEN_NS, SFQ_OR_GEN_NN, SFQ_OR_REC_X, SFQ_OR_USR_X, SFQ_OR_USR_XQ, SFQ_OR_USR_PFG, SFQ_OR_USR_PFQ,
                                      SFQ_OR_USR_LREC, SFQ_OR_USR_LGEN, SFQ_OR_USR_LQLT };
        for (int k = 0; k < NXS; k++) { out->data[sid[k]] = xs[k].rc.out.p; out->size[sid[k]] = xs[k].rc.out.n; xs[k].rc.out.p = NULL; }
    }
done:
    free(rc_rec.out.p); free(rc_gen.out.p); free(rc_qlt.out.p);
    if (xs) for (int k = 0; k < NXS; k++) free(xs[k].rc.out.p);
    free(xs); free(Q.ranger); free(G.ranger); free(H.ranger);
    if (e.failed) sfq_oracle_free_chunk(out);
    return e.failed;
}

int sfq_oracle_decode(const sfq_or_chunk *in, uint8_t **outp, size_t *out_n, char *err) {
    errctx e = { err, 0 };
    if (err) err[0] = 0;
    *outp = NULL; *out_n = 0;
    int level = in->level > 4 ? 4 : in->level < 1 ? 1 : in->level;
    qmodel Q; gmodel G; hmodel H;
    obuf o = { 0, 0, 0 };
    xload *xl = (xload *)calloc(NXS, sizeof(xload));
    uint8_t *m_rec = (uint8_t *)malloc(2 * (MAX_ID_LLEN + 1)), *m_gen = (uint8_t *)malloc(MAX_GN_LLEN + 4),
            *m_qlt = (uint8_t *)malloc(MAX_GN_LLEN + 2);
    int okq = q_init(&Q, level), okg = g_init(&G, level), okh = h_init(&H);
    if (!xl || !m_rec || !m_gen || !m_qlt || !okq || !okg || !okh) { fail(&e, "out of memory"); goto done; }
    H.pre5 = in->version < 5;                                                           /* (an absent key reads 0, config.cpp:373) */
    rc_dec rc_rec, rc_gen, rc_qlt;
    dec_init(&rc_rec, in->data[SFQ_OR_REC], in->size[SFQ_OR_REC]);
    dec_init(&rc_gen, in->data[SFQ_OR_GEN], in->size[SFQ_OR_GEN]);
    dec_init(&rc_qlt, in->data[SFQ_OR_QLT], in->size[SFQ_OR_QLT]);
    for (int k = 0; k < NXS; k++) xl_open(&xl[k], in->data[SFQ_OR_GEN_NS + k], in->size[SFQ_OR_GEN_NS + k]);
    xload *x_ns = &xl[0], *x_nn = &xl[1], *x_rec = &xl[2], *x_llen = &xl[3], *x_qlen = &xl[4],
          *x_sgen = &xl[5], *x_sqlt = &xl[6], *x_lrec = &xl[7], *x_lgen = &xl[8], *x_lqlt = &xl[9];

    /* UsrLoad ctor, usrs.cpp:411-456; GenLoad ctor gens.cpp:164-190; RecLoad ctor recs.cpp:93-106 */
    size_t m_llen = (size_t)in->llen, m_qlen = m_llen;
    int solid = in->solid, two_id = in->two_id;
    const char *alphabet = solid ? "0123" : "ACGT";
    G.n_byte = in->n_byte ? (uint8_t)in->n_byte : 'N';
    uint64_t i_llen = xl_get(x_llen), i_qlen = xl_get(x_qlen), i_sgen = xl_get(x_sgen), i_sqlt = xl_get(x_sqlt);
    G.ns_index = xl_get(x_ns);
    G.nn_index = xl_get(x_nn);
    H.index = xl_get(x_rec);
    uint8_t *hb[2] = { m_rec, m_rec + MAX_ID_LLEN + 1 };
    int flip = 0;
    uint8_t pf_gen = 0, pf_qlt = 0;
    uint64_t recno = 0;
    uint64_t i_long = xl_get(x_lrec);                                                    /* usrs.cpp:444-452 */
    if (!in->num_records && !i_long) { fail(&e, "Zero records, what's going on?"); goto done; }
    for (;;) {
        recno++;
        /* update(), usrs.cpp:471-510 */
        while (i_long == recno) {                                                        /* an oversized record, verbatim */
            xload *src[4] = { x_lrec, x_lgen, x_lrec, x_lqlt };
            ob_put(&o, '@');
            for (int l = 0; l < 4; l++) {
                if (!src[l]->valid) { fail(&e, "oversized-record stream missing"); goto done; }
                for (uint64_t guard = 0;; guard++) {
                    const uint8_t c = xl_get_chr(src[l]);
                    ob_put(&o, c);
                    if (c == '\n') break;
                    if (guard > ((uint64_t)1 << 32)) { fail(&e, "corrupt oversized-record stream"); goto done; }
                }
            }
            i_long += xl_get(x_lrec);
            recno++;
        }
        if (i_llen == recno) { m_llen = (size_t)xl_get(x_llen); m_qlen = m_llen; i_llen += xl_get(x_llen); }
        if (i_qlen == recno) { m_qlen = (size_t)xl_get(x_qlen); i_qlen += xl_get(x_qlen); }
        else if (m_qlen != m_llen) m_qlen = m_llen;
        if (solid && i_sgen == recno) { pf_gen = xl_get_chr(x_sgen); i_sgen += xl_get(x_sgen); }
        if (solid && i_sqlt == recno) { pf_qlt = xl_get_chr(x_sqlt); i_sqlt += xl_get(x_sqlt); }
        if (recno > in->num_records) break;
        if (m_llen >= MAX_GN_LLEN || m_qlen >= MAX_GN_LLEN) { fail(&e, "corrupt length stream"); break; }

        uint8_t *b_rec = hb[flip], *p_rec = hb[!flip];
        size_t rsz = h_load(&H, &rc_rec, x_rec, recno, b_rec, p_rec, in, &e);
        if (e.failed) break;
        if (!rsz) { fail(&e, "premature EOF - %llu records left", (unsigned long long)in->num_records + 1); break; }
        b_rec[rsz] = '\n';   /* the next record's tokeniser stops here, as on the reference's putline'd buffer */
        q_load(&Q, &rc_qlt, m_qlt, m_qlen);
        g_load(&G, &rc_gen, x_ns, x_nn, alphabet, m_gen, m_qlt, m_llen, m_qlen);
        /* save(), usrs.cpp:512-529 */
        ob_put(&o, '@');
        for (size_t k = 0; k < rsz; k++) ob_put(&o, b_rec[k]);
        ob_put(&o, '\n');
        if (solid) ob_put(&o, pf_gen);
        for (size_t k = 0; k < m_llen; k++) ob_put(&o, m_gen[k]);
        ob_put(&o, '\n');
        ob_put(&o, '+');
        if (two_id) for (size_t k = 0; k < rsz; k++) ob_put(&o, b_rec[k]);
        ob_put(&o, '\n');
        if (solid) ob_put(&o, pf_qlt);
        for (size_t k = 0; k < m_qlen; k++) ob_put(&o, m_qlt[k]);
        ob_put(&o, '\n');
        flip = !flip;
    }
done:
    free(xl); free(m_rec); free(m_gen); free(m_qlt); free(Q.ranger); free(G.ranger); free(H.ranger);
    if (e.failed) { free(o.p); return 1; }
    *outp = o.p; *out_n = o.n;
    return 0;
}

void sfq_oracle_free_chunk(sfq_or_chunk *c) {
    for (int k = 0; k < SFQ_OR_NSTREAMS; k++) { free(c->data[k]); c->data[k] = NULL; c->size[k] = 0; }
}
void sfq_oracle_free(void *p) { free(p); }
