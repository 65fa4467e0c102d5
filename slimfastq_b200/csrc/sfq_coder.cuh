// Range coder + the three adaptive frequency models, in the layouts the sm_100a kernels use.
//
// Behaviour follows the reference (coder.hpp, base2_ranger.hpp, log64_ranger.hpp,
// power_ranger.hpp); the data layouts do not:
//   * a 64/256-symbol model is NSYM packed 32-bit slots  [aux:8 | sym^index:8 | freq:16].
//     Storing sym XOR slot-index makes all-zero memory the reference's start state (every
//     unused slot k implicitly holds symbol k with freq 0, log64_ranger.hpp:103-105,126-127),
//     and the aux bytes of slots 0..5 hold total (24 bit), count and iend, so the 16 bytes of
//     slots 0..3 are the whole hot state of a context: one 16-byte load and one 16-byte store
//     per coded symbol in the common case (hot symbols migrate to the front slots).
//   * base contexts live in a per-chunk open-addressing hash table (or a direct table at
//     level 1) because a chunk can touch at most `nbases` of the 2^22..2^26 contexts and an
//     untouched context is by definition (3,3,3,3) (base2_ranger.hpp:69-72).
#pragma once
#include "sfq_common.cuh"
#include <string.h>

#define SFQ_RC_TOP (1u << 24)

// ------------------------------------------------------------------ 16-byte accesses / prefetch
struct SfqU4 { uint32_t x, y, z, w; };
SFQ_HD SfqU4 sfq_ld16(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    SfqU4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    SfqU4 r; memcpy(&r, p, 16); return r;
#endif
}
SFQ_HD void sfq_st16(uint32_t *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(p) = make_uint4(x, y, z, w);
#else
    p[0] = x; p[1] = y; p[2] = z; p[3] = w;
#endif
}
// Bring the sector holding *p towards the SM ahead of its use (contexts of an encoder are known
// from the input alone, so a run-ahead cursor can name them long before the coder needs them).
SFQ_HD void sfq_prefetch(const void *p) {
#if defined(__CUDA_ARCH__) && defined(SFQ_PREFETCH_L1)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif defined(__CUDA_ARCH__) && !defined(SFQ_NO_PREFETCH)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

SFQ_HD void sfq_prefetch_l1(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ------------------------------------------------------------------ byte sink / source
// Output bytes are gathered in a 64-bit register and stored 8 at a time (the arena sub-ranges are
// 16-byte aligned and their capacities multiples of 8).
struct SfqSink {
    uint8_t *p;
    uint32_t n, cap;
    uint64_t acc;
    bool mute;             // lane-cooperative coders: every lane counts, only one lane stores
    SFQ_HD void init(uint8_t *buf, uint32_t capacity) { p = buf; n = 0; cap = capacity & ~7u; acc = 0; mute = false; }
    SFQ_HD void put(uint8_t c) {
        acc |= (uint64_t)c << (8 * (n & 7u));
        n++;
        if ((n & 7u) == 0) {
            if (n <= cap && !mute) *reinterpret_cast<uint64_t *>(p + n - 8) = acc;
            acc = 0;
        }
    }
    SFQ_HD void flush() {                 // the last partial word; n > cap afterwards <=> overflow (SFQ_E_CAP)
        const uint32_t base = n & ~7u;
        for (uint32_t k = base; k < n; k++)
            if (k < cap && !mute) p[k] = (uint8_t)(acc >> (8 * (k - base)));
    }
    SFQ_HD bool overflow() const { return n > cap; }
};

// Sequential reader over text bytes: one aligned 8-byte load per 8 symbols.
struct SfqReader {
    const uint8_t *p;
    uint64_t word;
    SFQ_HD void seek(const uint8_t *q) {
        p = q;
        word = *reinterpret_cast<const uint64_t *>((uintptr_t)q & ~(uintptr_t)7);
    }
    SFQ_HD uint8_t next() {
        const uint32_t k = (uint32_t)((uintptr_t)p & 7u);
        if (k == 0) word = *reinterpret_cast<const uint64_t *>(p);
        p++;
        return (uint8_t)(word >> (8 * k));
    }
};

// ------------------------------------------------------------------ encoder  (coder.hpp:34-39,52-81)
struct SfqEnc {
    uint64_t low;
    uint32_t range;
    SfqSink out;
    bool live;
    SFQ_HD void reset() { live = false; low = 0; range = 0xFFFFFFFFu; out.p = nullptr; out.n = 0; out.cap = 0; out.acc = 0; out.mute = false; }
    SFQ_HD void start(uint8_t *buf, uint32_t cap) { low = 0; range = 0xFFFFFFFFu; out.init(buf, cap); live = true; }
    SFQ_HD void encode(uint32_t cum, uint32_t freq, uint32_t tot) {
        range /= tot;
        low += (uint32_t)(cum * range);
        range *= freq;
        while (range < SFQ_RC_TOP) {
            if ((low ^ (low + range)) & (0xffULL << 56))
                range = (((uint32_t)low) | (SFQ_RC_TOP - 1)) - (uint32_t)low;
            out.put((uint8_t)(low >> 56));
            range <<= 8;
            low <<= 8;
        }
    }
    // encode() with the quotient r = range / totFreq supplied by the caller (reciprocal multiply)
    SFQ_HD void encode_scaled(uint32_t cum, uint32_t freq, uint32_t r) {
        low += (uint32_t)(cum * r);
        range = r * freq;
        while (range < SFQ_RC_TOP) {
            if ((low ^ (low + range)) & (0xffULL << 56))
                range = (((uint32_t)low) | (SFQ_RC_TOP - 1)) - (uint32_t)low;
            out.put((uint8_t)(low >> 56));
            range <<= 8;
            low <<= 8;
        }
    }
    SFQ_HD void finish() {
        for (int i = 0; i < 8; i++) { out.put((uint8_t)(low >> 56)); low <<= 8; }
        out.flush();
    }
};

// ------------------------------------------------------------------ decoder  (coder.hpp:41-49,83-102)
struct SfqDec {
    uint64_t low, code;
    uint32_t range;
    const uint8_t *p;
    uint32_t n, pos;
    uint64_t word;         // the aligned 8 bytes of the stream that hold byte `pos`
    bool valid, have;
    SFQ_HD uint8_t next() {                                             // EOF reads as 0, filer.hpp:94-97
        if (pos >= n) return 0;
        const uint8_t *q = p + pos;
        const uint32_t k = (uint32_t)((uintptr_t)q & 7u);
        if (k == 0 || !have) { word = *reinterpret_cast<const uint64_t *>(q - k); have = true; }
        pos++;
        return (uint8_t)(word >> (8 * k));
    }
    SFQ_HD void start(const uint8_t *buf, uint32_t size) {
        p = buf; n = size; pos = 0; low = 0; range = 0xFFFFFFFFu; code = 0; word = 0; have = false;
        valid = (buf != nullptr && size != 0);
        for (int i = 0; i < 8; i++) code = (code << 8) | next();
    }
    SFQ_HD uint32_t get_freq(uint32_t tot) {
        range /= tot;
        // code < range * tot <= 2^32 holds for any stream an encoder can write; the 64-bit form is
        // kept so that a corrupt stream behaves like the reference instead of trapping.
        return (code >> 32) ? (uint32_t)(code / range) : ((uint32_t)code / range);
    }
    SFQ_HD void decode(uint32_t cum, uint32_t freq) {
        uint32_t t = cum * range;
        low += t;
        code -= t;
        range *= freq;
        while (range < SFQ_RC_TOP) {
            if ((low ^ (low + range)) & (0xffULL << 56))
                range = (((uint32_t)low) | (SFQ_RC_TOP - 1)) - (uint32_t)low;
            code = (code << 8) | next();
            range <<= 8;
            low <<= 8;
        }
    }
};

// ------------------------------------------------------------------ 4-symbol model (base2_ranger.hpp:35-105)
// v holds freq[0..3] in bytes 0..3.
SFQ_HD uint32_t sfq_b2_update(uint32_t v, uint32_t s) {
    if (((v >> (8 * s)) & 0xffu) > 254u)
        v = ((v & 0xFEFEFEFEu) >> 1) | (v & 0x01010101u);
    return v + (1u << (8 * s));
}
SFQ_HD uint32_t sfq_b2_put(uint32_t v, SfqEnc &rc, uint32_t s) {
    uint32_t f0 = v & 0xff, f1 = (v >> 8) & 0xff, f2 = (v >> 16) & 0xff, f3 = v >> 24;
    uint32_t tot = f0 + f1 + f2 + f3;
    uint32_t cum = (s > 0 ? f0 : 0) + (s > 1 ? f1 : 0) + (s > 2 ? f2 : 0);
    uint32_t f = (v >> (8 * s)) & 0xff;
    rc.encode(cum, f, tot);
    return sfq_b2_update(v, s);
}
SFQ_HD uint32_t sfq_b2_get(uint32_t v, SfqDec &rc, uint32_t &sym) {
    uint32_t f0 = v & 0xff, f1 = (v >> 8) & 0xff, f2 = (v >> 16) & 0xff, f3 = v >> 24;
    uint32_t tot = f0 + f1 + f2 + f3;
    uint32_t prob = rc.get_freq(tot);
    uint32_t s, cum, f;
    if (prob < f0) { s = 0; cum = 0; f = f0; }
    else if (prob < f0 + f1) { s = 1; cum = f0; f = f1; }
    else if (prob < f0 + f1 + f2) { s = 2; cum = f0 + f1; f = f2; }
    else { s = 3; cum = f0 + f1 + f2; f = f3; }
    rc.decode(cum, f);
    sym = s;
    return sfq_b2_update(v, s);
}

// Per-chunk base-context table.  `dense` (level 1: 2^18 contexts) stores freq ^ 0x03030303 so
// zeroed memory is the start state; otherwise 64-bit slots (ctx+1)<<32 | freq with linear probing.
struct SfqGenTable {
    uint64_t *slots;       // hash slots, or the dense u32 table reinterpret-cast
    uint32_t hbits;        // log2(#hash slots)
    uint32_t used;         // occupied hash slots
    uint32_t dense;
    SFQ_HD void init(void *mem, uint32_t hash_bits, bool is_dense) {
        slots = (uint64_t *)mem; hbits = hash_bits; used = 0; dense = is_dense;
    }
    // Returns the slot index of ctx (inserting the start state if absent) or 0xFFFFFFFF if full.
    SFQ_HD uint32_t find(uint32_t ctx, uint32_t &v) {
        if (dense) {
            v = ((const uint32_t *)slots)[ctx] ^ 0x03030303u;
            return ctx;
        }
        const uint32_t h = (ctx * 2654435761u) >> (32 - hbits);
        const uint64_t k = slots[h];
        const uint32_t kk = (uint32_t)(k >> 32);
        if (kk == ctx + 1u) { v = (uint32_t)k; return h; }          // the usual case: home slot
        return find_probe(ctx, h, kk, v);
    }
    SFQ_COLD uint32_t find_probe(uint32_t ctx, uint32_t h, uint32_t kk, uint32_t &v) {
        const uint32_t mask = (1u << hbits) - 1u;
        const uint32_t key = ctx + 1u;
        for (;;) {
            if (kk == 0) {
                if (used + 1u >= mask) return 0xFFFFFFFFu;
                used++;
                v = 0x03030303u;
                return h;
            }
            h = (h + 1u) & mask;
            const uint64_t k = slots[h];
            kk = (uint32_t)(k >> 32);
            if (kk == key) { v = (uint32_t)k; return h; }
        }
    }
    SFQ_HD void prefetch(uint32_t ctx) const {
        if (dense) sfq_prefetch((const uint32_t *)slots + ctx);
        else sfq_prefetch(slots + ((ctx * 2654435761u) >> (32 - hbits)));
    }
    SFQ_HD void store(uint32_t slot, uint32_t ctx, uint32_t v) {
        if (dense) ((uint32_t *)slots)[slot] = v ^ 0x03030303u;
        else slots[slot] = ((uint64_t)(ctx + 1u) << 32) | v;
    }
};

// ------------------------------------------------------------------ 64 / 256-symbol models
// log64_ranger.hpp:36-140 (NSYM 64, STEP 6, MAX_FREQ 65472, saturation slack 20) and
// power_ranger.hpp:36-131 (256, 14, 32736, 256) are one scheme; see the layout note on top.
template <int NSYM, int STEP, int MAXF, int SLACK>
struct SfqAModel {
    uint32_t *m;   // NSYM packed slots in global memory

    SFQ_HD static uint32_t freq_of(uint32_t s) { return s & 0xffffu; }
    SFQ_HD static uint32_t sym_of(uint32_t s, uint32_t i) { return ((s >> 16) & 0xffu) ^ (i & 0xffu); }
    SFQ_HD static uint32_t aux_of(uint32_t s) { return s >> 24; }
    SFQ_HD static uint32_t pack(uint32_t aux, uint32_t sym, uint32_t i, uint32_t f) {
        return (aux << 24) | (((sym ^ i) & 0xffu) << 16) | f;
    }
    SFQ_HD uint32_t total() const { return aux_of(m[0]) | (aux_of(m[1]) << 8) | (aux_of(m[2]) << 16); }
    SFQ_HD void set_total(uint32_t t) {
        m[0] = (m[0] & 0x00ffffffu) | ((t & 0xffu) << 24);
        m[1] = (m[1] & 0x00ffffffu) | (((t >> 8) & 0xffu) << 24);
        m[2] = (m[2] & 0x00ffffffu) | (((t >> 16) & 0xffu) << 24);
    }
    SFQ_HD uint32_t iend() const { return aux_of(m[4]) | (aux_of(m[5]) << 8); }
    SFQ_HD void set_iend(uint32_t e) {
        m[4] = (m[4] & 0x00ffffffu) | ((e & 0xffu) << 24);
        m[5] = (m[5] & 0x00ffffffu) | (((e >> 8) & 0xffu) << 24);
    }

    // update_freq (log64:69-87, power:68-86).  Returns the symbol that was in slot i.
    SFQ_HD uint32_t update(uint32_t i, uint32_t tot) {
        uint32_t s = m[i];
        uint32_t f = freq_of(s);
        const uint32_t sym = sym_of(s, i);
        if (f > (uint32_t)(MAXF - STEP)) {
            if (i == 0 && f + (uint32_t)SLACK > tot) return sym;
            const uint32_t e = iend();
            tot = 0;
            for (uint32_t j = 0; j < e; j++) {       // normalize(): halve every live slot
                uint32_t sj = m[j];
                uint32_t fj = freq_of(sj) >> 1;
                m[j] = (sj & 0xffff0000u) | fj;
                tot += fj;
            }
            s = m[i];
            f = freq_of(s);
        }
        f += STEP;
        tot += STEP;
        m[i] = (s & 0xffff0000u) | f;
        set_total(tot);
        if (i == 0) return sym;                      // `++count` is not evaluated for slot 0
        uint32_t c3 = m[3];
        uint32_t count = (aux_of(c3) + 1u) & 0xffu;
        m[3] = (c3 & 0x00ffffffu) | (count << 24);
        if ((count & 0xfu) == 0) {
            uint32_t a = m[i], b = m[i - 1];
            if (freq_of(a) > freq_of(b)) {           // down_level(): swap slots i and i-1
                uint32_t sa = sym_of(a, i), sb = sym_of(b, i - 1);
                m[i]     = pack(aux_of(a), sb, i,     freq_of(b));
                m[i - 1] = pack(aux_of(b), sa, i - 1, freq_of(a));
            }
        }
        return sym;
    }

    // Reference-shaped paths over memory: any slot, normalisation, corrupt input.
    template <class RC> SFQ_COLD void put_slow(RC &rc, uint32_t sym) {      // log64:98-112, power:93-106
        if (iend() <= sym) set_iend(sym + 1u);
        uint32_t i = 0, sumf = 0, s;
        for (;; i++) {
            s = m[i];
            if (sym_of(s, i) == sym) break;
            sumf += freq_of(s);
        }
        const uint32_t tot = total();
        rc.encode(sumf + i, freq_of(s) + 1u, tot + NSYM);
        update(i, tot);
    }
    SFQ_COLD uint32_t get_slow(SfqDec &rc, uint32_t prob, uint32_t tot) {   // log64:114-138, power:108-130
        uint32_t i = 0, sumf = 0, f = 0;
        for (; i < NSYM; i++) {
            f = freq_of(m[i]);
            if (sumf + f + 1u <= prob) sumf += f + 1u; else break;
        }
        if (i >= NSYM) { i = NSYM - 1; f = freq_of(m[i]); }   // only a corrupt stream gets here
        if (iend() <= i) set_iend(i + 1u);
        rc.decode(sumf, f + 1u);
        return update(i, tot);
    }

    // Register paths: the symbol sits in slots 0..7 (one 32-byte sector) and no halving is due -
    // nearly always, since hot symbols migrate to the front.  Everything happens in registers
    // between one pair of 16-byte loads and one or two 16-byte stores.
    SFQ_HD static void set_aux(uint32_t &s, uint32_t a) { s = (s & 0x00ffffffu) | (a << 24); }
    SFQ_HD void finish_fast(uint32_t (&w)[8], uint32_t i, uint32_t f, uint32_t tot, uint32_t iend_old, uint32_t iend_new) {
        f += STEP;
        tot += STEP;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) if (k == i) w[k] = (w[k] & 0xffff0000u) | f;
        set_aux(w[0], tot & 0xffu); set_aux(w[1], (tot >> 8) & 0xffu); set_aux(w[2], (tot >> 16) & 0xffu);
        if (iend_new != iend_old) { set_aux(w[4], iend_new & 0xffu); set_aux(w[5], (iend_new >> 8) & 0xffu); }
        if (i != 0) {
            const uint32_t count = (aux_of(w[3]) + 1u) & 0xffu;
            set_aux(w[3], count);
            if ((count & 0xfu) == 0) {
#pragma unroll
                for (uint32_t k = 1; k < 8; k++)
                    if (k == i) {
                        const uint32_t a = w[k], b = w[k - 1];
                        if (freq_of(a) > freq_of(b)) {
                            w[k]     = pack(aux_of(a), sym_of(b, k - 1), k,     freq_of(b));
                            w[k - 1] = pack(aux_of(b), sym_of(a, k),     k - 1, freq_of(a));
                        }
                    }
            }
        }
        sfq_st16(m, w[0], w[1], w[2], w[3]);
        if (i >= 4 || iend_new != iend_old) sfq_st16(m + 4, w[4], w[5], w[6], w[7]);
    }

    template <class RC> SFQ_HD void put(RC &rc, uint32_t sym) {
        uint32_t w[8];
        { const SfqU4 a = sfq_ld16(m), b = sfq_ld16(m + 4);
          w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; }
        uint32_t i = 8, sumf = 0, f = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++)
            if (i == 8) { if (sym_of(w[k], k) == sym) { i = k; f = freq_of(w[k]); } else sumf += freq_of(w[k]); }
        if (i == 8 || f > (uint32_t)(MAXF - STEP)) { put_slow(rc, sym); return; }
        const uint32_t tot = aux_of(w[0]) | (aux_of(w[1]) << 8) | (aux_of(w[2]) << 16);
        const uint32_t ie = aux_of(w[4]) | (aux_of(w[5]) << 8);
        rc.encode(sumf + i, f + 1u, tot + NSYM);
        finish_fast(w, i, f, tot, ie, ie <= sym ? sym + 1u : ie);
    }

    SFQ_HD uint32_t get(SfqDec &rc) {
        uint32_t w[8];
        { const SfqU4 a = sfq_ld16(m), b = sfq_ld16(m + 4);
          w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; }
        const uint32_t tot = aux_of(w[0]) | (aux_of(w[1]) << 8) | (aux_of(w[2]) << 16);
        const uint32_t prob = rc.get_freq(tot + NSYM);
        uint32_t i = 8, sumf = 0, f = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++)
            if (i == 8) { const uint32_t fk = freq_of(w[k]); if (sumf + fk + 1u <= prob) sumf += fk + 1u; else { i = k; f = fk; } }
        if (i == 8 || f > (uint32_t)(MAXF - STEP)) return get_slow(rc, prob, tot);
        uint32_t sym = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) if (k == i) sym = sym_of(w[k], k);
        const uint32_t ie = aux_of(w[4]) | (aux_of(w[5]) << 8);
        rc.decode(sumf, f + 1u);
        finish_fast(w, i, f, tot, ie, ie <= i ? i + 1u : ie);
        return sym;
    }
};
typedef SfqAModel<64, 6, 65472, 20> SfqLog64;

// PowerRanger (power_ranger.hpp:36-131: 256 symbols, STEP 14, MAX_FREQ 32736) for one thread, in a
// layout that needs no scan: the reference finds a symbol's slot and its cumulative frequency by
// walking the slots (on average 128 of them for header text and number bytes); here an inverse map
// gives the slot and per-16-slot group sums give the cumulative frequency in <= 15 + 15 additions.
// Zeroed memory is the start state (symbols stored XOR slot index, positions XOR symbol).
//   words   0..255  slots  [sym ^ index : 8 << 16 | freq : 16]
//   words 256..319  inverse map, one byte per symbol: slot ^ symbol
//   words 320..335  gsum[16]: sum of freq over slots 16g..16g+15
//   word  336       total          word 337  count (power_ranger.hpp:43-45)
struct SfqPower {
    uint32_t *m;
    SFQ_HD static uint32_t freq_of(uint32_t s) { return s & 0xffffu; }
    SFQ_HD static uint32_t sym_of(uint32_t s, uint32_t i) { return ((s >> 16) & 0xffu) ^ (i & 0xffu); }
    SFQ_HD static uint32_t pack(uint32_t sym, uint32_t i, uint32_t f) { return (((sym ^ i) & 0xffu) << 16) | f; }
    SFQ_HD uint8_t *inv() const { return reinterpret_cast<uint8_t *>(m + 256); }
    SFQ_HD uint32_t *gsum() const { return m + 320; }

    // update_freq (power_ranger.hpp:68-86) for slot i holding `sym` with frequency f
    SFQ_HD void update(uint32_t i, uint32_t f, uint32_t sym, uint32_t tot) {
        if (f > 32736u - 14u) {
            if (i == 0 && f + 256u > tot) return;
            tot = 0;
            for (uint32_t g = 0; g < 16; g++) {                    // normalize(): halve every slot
                uint32_t gs = 0;
                for (uint32_t k = 16 * g; k < 16 * g + 16; k++) { const uint32_t sk = m[k], fk = freq_of(sk) >> 1; m[k] = (sk & 0xffff0000u) | fk; gs += fk; }
                gsum()[g] = gs;
                tot += gs;
            }
            f = freq_of(m[i]);
        }
        f += 14u;
        m[i] = pack(sym, i, f);
        gsum()[i >> 4] += 14u;
        m[336] = tot + 14u;
        if (i == 0) return;                                        // `++count` is not evaluated for slot 0
        const uint32_t count = (m[337] + 1u) & 0xffu;
        m[337] = count;
        if ((count & 0xfu) == 0) {
            const uint32_t b = m[i - 1], fb = freq_of(b);
            if (f > fb) {                                          // swap slots i and i-1
                const uint32_t symb = sym_of(b, i - 1);
                m[i] = pack(symb, i, fb);
                m[i - 1] = pack(sym, i - 1, f);
                inv()[sym] = (uint8_t)((i - 1) ^ sym);
                inv()[symb] = (uint8_t)(i ^ symb);
                if ((i & 15u) == 0) { gsum()[(i >> 4) - 1] += f - fb; gsum()[i >> 4] -= f - fb; }
            }
        }
    }
    // The sixteen group sums and the sixteen slots of one group, each as four independent 16-byte loads: the walk over
    // them then runs in registers.  (Read one word at a time in a loop, every step of the walk waited for its own L2 round
    // trip: ~30 of them per coded symbol, which is what the header coders spent their time on.)
    SFQ_HD void load16(const uint32_t *p, uint32_t (&w)[16]) const {
#pragma unroll
        for (int q = 0; q < 4; q++) { const SfqU4 v = sfq_ld16(p + 4 * q); w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
    }
    template <class RC> SFQ_HD void put(RC &rc, uint32_t sym) {    // power_ranger.hpp:93-106
        const uint32_t i = (uint32_t)inv()[sym] ^ sym, g = i >> 4, k = i & 15u;
        uint32_t gs[16], sl[16];
        load16(gsum(), gs);
        load16(m + 16u * g, sl);
        const uint32_t tot = m[336];
        uint32_t sumf = 0, f = 0;
#pragma unroll
        for (uint32_t j = 0; j < 16; j++) {
            sumf += j < g ? gs[j] : 0u;
            sumf += j < k ? freq_of(sl[j]) : 0u;
            f = j == k ? freq_of(sl[j]) : f;
        }
        rc.encode(sumf + i, f + 1u, tot + 256u);
        update(i, f, sym, tot);
    }
    template <class RC> SFQ_HD uint32_t get(RC &rc) {              // power_ranger.hpp:108-130
        const uint32_t tot = m[336];
        uint32_t gs[16];
        load16(gsum(), gs);
        const uint32_t prob = rc.get_freq(tot + 256u);
        uint32_t cum = 0, g = 0;
        bool go = true;
#pragma unroll
        for (uint32_t j = 0; j < 15; j++) {                        // the last group also catches a corrupt stream
            const uint32_t t = gs[j] + 16u;
            go = go && cum + t <= prob;
            cum += go ? t : 0u; g += go ? 1u : 0u;
        }
        uint32_t sl[16];
        load16(m + 16u * g, sl);
        uint32_t k = 0, f = freq_of(sl[0]), word = sl[0];
        go = true;
#pragma unroll
        for (uint32_t j = 0; j < 15; j++) {                        // the last slot also catches a corrupt stream
            const uint32_t fj = freq_of(sl[j]);
            go = go && cum + fj + 1u <= prob;
            cum += go ? fj + 1u : 0u; k += go ? 1u : 0u;
            f = go ? freq_of(sl[j + 1]) : f; word = go ? sl[j + 1] : word;
        }
        const uint32_t i = 16u * g + k;
        const uint32_t sym = sym_of(word, i);
        rc.decode(cum, f + 1u);
        update(i, f, sym, tot);
        return sym;
    }
};

// PowerRangerU: variable-length u64 over 14 consecutive 256-symbol models (power_ranger.hpp:133-192)
struct SfqPowerU {
    uint32_t *m;   // 14 models, SFQ_PW_WORDS apart
    SFQ_HD SfqPower at(int k) const { SfqPower p; p.m = m + (size_t)k * SFQ_PW_WORDS; return p; }
    SFQ_COLD void put(SfqEnc &rc, uint64_t num) {
        if (num <= 0x7f) { at(0).put(rc, (uint32_t)num); return; }
        if (num < 0x7ffe) {
            at(0).put(rc, (uint32_t)(0xff & (0x80 | (num >> 8))));
            at(1).put(rc, (uint32_t)(0xff & num));
            return;
        }
        at(0).put(rc, 0xff);
        if (num < (1ULL << 32)) {
            at(1).put(rc, 0xfe);
            for (int sh = 0, i = 2; sh < 32; sh += 8, i++) at(i).put(rc, (uint32_t)(0xff & (num >> sh)));
        } else {
            at(1).put(rc, 0xff);
            for (int sh = 0, i = 6; sh < 64; sh += 8, i++) at(i).put(rc, (uint32_t)(0xff & (num >> sh)));
        }
    }
    SFQ_COLD uint64_t get(SfqDec &rc) {
        uint64_t num = at(0).get(rc);
        if (num > 0x7f) {
            num = (num << 8) | at(1).get(rc);
            if (num < 0xfffe) num &= 0x7fff;
            else if (num == 0xfffe) {
                num = 0;
                for (int sh = 0, i = 2; sh < 32; sh += 8, i++) num |= (uint64_t)at(i).get(rc) << sh;
            } else {
                num = 0;
                for (int sh = 0, i = 6; sh < 64; sh += 8, i++) num |= (uint64_t)at(i).get(rc) << sh;
            }
        }
        return num;
    }
};

// ------------------------------------------------------------------ exception-list streams (xfile.cpp:36-110)
// Lazily created on the first put; closing writes put(0) and the coder flush.
struct SfqXSave {
    SfqEnc rc;
    SfqPowerU num;
    SfqPower str;
    uint8_t *buf;
    uint32_t cap;
    SFQ_HD void init(uint32_t *pool, int xslot, uint8_t *out, uint32_t capacity) {
        rc.reset();
        num.m = pool + (size_t)(SFQ_PW_X_BASE + xslot * SFQ_PW_PER_X) * SFQ_PW_WORDS;
        str.m = num.m + (size_t)14 * SFQ_PW_WORDS;
        buf = out; cap = capacity;
    }
    SFQ_HD void open() { if (!rc.live) rc.start(buf, cap); }
    SFQ_COLD void put(uint64_t v) { open(); num.put(rc, v); }
    SFQ_COLD void put_chr(uint8_t c) { open(); str.put(rc, c); }
    // returns stream size (0 = never created); sets ovf on arena overflow
    SFQ_COLD uint32_t close(bool &ovf) {
        if (!rc.live) return 0;
        put(0);
        rc.finish();
        if (rc.out.overflow()) ovf = true;
        return rc.out.n;
    }
};
struct SfqXLoad {
    SfqDec rc;
    SfqPowerU num;
    SfqPower str;
    SFQ_HD void init(uint32_t *pool, int xslot, const uint8_t *in, uint32_t size) {
        num.m = pool + (size_t)(SFQ_PW_X_BASE + xslot * SFQ_PW_PER_X) * SFQ_PW_WORDS;
        str.m = num.m + (size_t)14 * SFQ_PW_WORDS;
        rc.start(in, size);
    }
    SFQ_COLD uint64_t get() { return rc.valid ? num.get(rc) : 0; }   // absent stream reads 0 forever, xfile.cpp:90-93
    SFQ_COLD uint8_t get_chr() { return (uint8_t)str.get(rc); }
};
