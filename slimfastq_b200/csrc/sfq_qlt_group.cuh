// Quality stream, lane-cooperative: SFQ_QG lanes of a warp drive ONE adaptive coder.
//
// Inside a 1 MiB chunk a quality context is visited ~100 times, far too few for the reference's
// swap-every-16th-update ordering (log64_ranger.hpp:82-86) to bring hot symbols to the front: the
// coded symbol usually sits near its identity slot (30..40), and the reference's linear scan
// (log64_ranger.hpp:107, :120-130) is what the serial chain spends its time in.  Here the 64 slots
// of a context are spread over 8 lanes (8 slots = one 32-byte sector each): one coalesced 256-byte
// load, an 8-way local compare, a warp ballot to find the slot and a 3-step shuffle scan for the
// cumulative frequency replace the scan.  The range-coder arithmetic is computed redundantly by the
// 8 lanes (uniform), lane 0 of the group owns the output bytes.
//
// Behaviour is identical to SfqLog64::put_slow/get_slow (sfq_coder.cuh), which the CPU emulation
// checks against the oracle; this file is checked on the GPU (tests/test_gpu_parity.py).
#pragma once
#include "sfq_streams.cuh"

#if defined(__CUDACC__)

#define SFQ_QG 8                       // lanes per chunk-stream
#define SFQ_QS (64 / SFQ_QG)           // slots per lane

// Quality contexts of one chunk: direct table (level 1, or when the hash would be as large), else an
// open-addressing hash of 256-byte models keyed by the context (key in the aux bytes of slots 6,7,
// "occupied" in bit 7 of slot 5's aux byte; all of them live in lane 0's sector).
struct SfqQTable {
    uint32_t *base;
    uint32_t nent, used, dense;          // nent = entries of the hash (any size, not only powers of two)
    __device__ __forceinline__ void init(uint32_t *mem, uint32_t entries, bool is_dense) { base = mem; nent = entries; used = 0; dense = is_dense; }
    __device__ __forceinline__ uint32_t home(uint32_t ctx) const { return dense ? ctx : __umulhi(ctx * 2654435761u, nent); }
    __device__ __forceinline__ void prefetch(uint32_t ctx, uint32_t lane) const { sfq_prefetch(base + (size_t)home(ctx) * 64 + lane * SFQ_QS); }
};

// Run-ahead cursor of the quality encoder: replays the context function on the input SFQ_QLT_AHEAD
// symbols ahead of the coder (contexts depend on the input alone) and prefetches the sector of the
// context's home entry that this lane will load.
struct SfqQltCursor {
    const uint8_t *text; const uint64_t *ls; uint64_t line0;
    uint32_t nrec, solid, r, i, qlen;
    int level;
    SfqQCtx c;
    SfqReader rd;
    __device__ __forceinline__ void open_record() {
        while (r < nrec) {
            const SfqRecView v = sfq_rec_view(text, ls, line0, r, solid);
            qlen = v.qlen; i = 0; c.reset();
            if (qlen) { rd.seek(v.qual); return; }
            r++;
        }
        qlen = 0;
    }
    __device__ __forceinline__ void init(const uint8_t *t, const uint64_t *l, const SfqChunkMeta *m, int lvl) {
        text = t; ls = l; line0 = m->line0; nrec = m->nrec; solid = m->solid; level = lvl; r = 0;
        open_record();
    }
    __device__ __forceinline__ void run(const SfqQTable &tab, uint32_t lane, uint32_t count) {
        while (count && r < nrec) {
            tab.prefetch(c.last, lane);
            sfq_q_next(c, level, (uint8_t)(rd.next() - '!'));
            count--;
            if (++i == qlen) { r++; open_record(); }
        }
    }
};

struct SfqQGroup {
    unsigned gmask;      // lanes of this group within the warp
    uint32_t lane;       // 0..7 inside the group
    uint32_t gbase;      // first lane of the group inside the warp
    uint32_t w[SFQ_QS];  // this lane's 8 slots of the current context
    uint32_t *m;         // current context (group-uniform)

    __device__ __forceinline__ uint32_t bcast(uint32_t v, uint32_t src) const { return __shfl_sync(gmask, v, src, SFQ_QG); }
    __device__ __forceinline__ void load(uint32_t *model) {
        m = model;
        const uint4 a = *reinterpret_cast<const uint4 *>(model + lane * SFQ_QS);
        const uint4 b = *reinterpret_cast<const uint4 *>(model + lane * SFQ_QS + 4);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
    __device__ __forceinline__ void store() const {
        *reinterpret_cast<uint4 *>(m + lane * SFQ_QS) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4 *>(m + lane * SFQ_QS + 4) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    // Finds (or claims) the model of `ctx`, leaves it loaded.  False = table full.
    __device__ __forceinline__ bool locate(SfqQTable &t, uint32_t ctx) {
        if (t.dense) { load(t.base + (size_t)ctx * 64); return true; }
        uint32_t h = t.home(ctx);
        for (;;) {
            load(t.base + (size_t)h * 64);
            const uint32_t a5 = bcast(w[5] >> 24, 0);
            const uint32_t key = bcast((w[6] >> 24) | ((w[7] >> 24) << 8), 0);
            if (a5 & 0x80u) { if (key == ctx) return true; }
            else {
                if (t.used + 1u >= t.nent) return false;
                t.used++;
                if (lane == 0) {
                    w[5] |= 0x80000000u;
                    w[6] = (w[6] & 0x00ffffffu) | ((ctx & 0xffu) << 24);
                    w[7] = (w[7] & 0x00ffffffu) | ((ctx >> 8) << 24);
                }
                return true;      // lane 0 is stored by every update, which persists the claim
            }
            h = h + 1u == t.nent ? 0u : h + 1u;
        }
    }

    // update_freq (log64_ranger.hpp:69-87) on slot i = 8*hl + hk holding frequency f; tot/count are the
    // context's current values.  Stores whatever changed.  (`iend` is not kept: with the sym^index slot
    // encoding every slot is always valid, and slots past the reference's iend have freq 0, so its only
    // uses - the bounds of the search and of normalize() - change nothing.)
    __device__ __forceinline__ void update(uint32_t hl, uint32_t hk, uint32_t f, uint32_t tot, uint32_t count) {
        const uint32_t i = hl * SFQ_QS + hk;
        bool all_dirty = false;
        if (f > 65472u - 6u) {
            if (i == 0 && f + 20u > tot) {                  // saturated front slot: no update at all
                if (lane == 0) store();                     // (persists a fresh claim)
                return;
            }
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < SFQ_QS; k++) { const uint32_t fk = (w[k] & 0xffffu) >> 1; w[k] = (w[k] & 0xffff0000u) | fk; s += fk; }
            tot = __reduce_add_sync(gmask, s);
            f >>= 1;
            all_dirty = true;
        }
        f += 6u;
        tot += 6u;
        if (lane == hl) {
#pragma unroll
            for (uint32_t k = 0; k < SFQ_QS; k++) if (k == hk) w[k] = (w[k] & 0xffff0000u) | f;
        }
        bool prev_dirty = false;
        if (i != 0) {
            count = (count + 1u) & 0xffu;
            if (lane == 0) w[3] = (w[3] & 0x00ffffffu) | (count << 24);
            if ((count & 0xfu) == 0) {                      // maybe swap slot i with slot i-1
                // the neighbour is slot hk-1 of the same lane, or slot 7 of the lane before
                const uint32_t left7 = __shfl_up_sync(gmask, w[SFQ_QS - 1], 1, SFQ_QG);
                uint32_t cur = 0, prv = 0;
#pragma unroll
                for (uint32_t k = 0; k < SFQ_QS; k++) if (k == hk) { cur = w[k]; prv = k ? w[k ? k - 1 : 0] : left7; }
                cur = bcast(cur, hl); prv = bcast(prv, hl);
                if ((cur & 0xffffu) > (prv & 0xffffu)) {
                    const uint32_t sym_i = ((cur >> 16) & 0xffu) ^ i, sym_p = ((prv >> 16) & 0xffu) ^ (i - 1u);
                    const uint32_t new_i = (cur & 0xff000000u) | (((sym_p ^ i) & 0xffu) << 16) | (prv & 0xffffu);
                    const uint32_t new_p = (prv & 0xff000000u) | (((sym_i ^ (i - 1u)) & 0xffu) << 16) | (cur & 0xffffu);
                    if (lane == hl) {
#pragma unroll
                        for (uint32_t k = 0; k < SFQ_QS; k++) {
                            if (k == hk) w[k] = new_i;
                            if (hk > 0 && k + 1 == hk) w[k] = new_p;
                        }
                    }
                    if (hk == 0 && lane + 1 == hl) { w[SFQ_QS - 1] = new_p; prev_dirty = true; }
                }
            }
        }
        if (lane == 0) {
            w[0] = (w[0] & 0x00ffffffu) | ((tot & 0xffu) << 24);
            w[1] = (w[1] & 0x00ffffffu) | (((tot >> 8) & 0xffu) << 24);
            w[2] = (w[2] & 0x00ffffffu) | (((tot >> 16) & 0xffu) << 24);
        }
        if (all_dirty || lane == 0 || lane == hl || prev_dirty) store();
    }
    // total (bits 0..23) and count (bits 24..31) of the loaded context, from lane 0
    __device__ __forceinline__ uint32_t header() const {
        return bcast((w[0] >> 24) | ((w[1] >> 24) << 8) | ((w[2] >> 24) << 16) | (w[3] & 0xff000000u), 0);
    }

    // Log64Ranger::put (log64_ranger.hpp:98-112) on the loaded context: 8-way local compare, one
    // ballot for the slot, two warp reductions for the cumulative frequency and the slot's own.
    __device__ __forceinline__ void put(SfqEnc &rc, uint32_t sym) {
        int hit = -1;
        uint32_t before = 0, lsum = 0, fh = 0;
#pragma unroll
        for (int k = 0; k < SFQ_QS; k++) {
            const uint32_t f = w[k] & 0xffffu;
            const uint32_t s = ((w[k] >> 16) & 0xffu) ^ (lane * SFQ_QS + k);
            if (hit < 0) { if (s == sym) { hit = k; fh = f; } else before += f; }
            lsum += f;
        }
        const uint32_t hdr = header();
        const unsigned ball = (__ballot_sync(gmask, hit >= 0) >> gbase) & 0xffu;
        const uint32_t hl = (uint32_t)(__ffs(ball) - 1);
        const uint32_t sumf = __reduce_add_sync(gmask, lane < hl ? lsum : (lane == hl ? before : 0u));
        const uint32_t fk = __reduce_or_sync(gmask, lane == hl ? (fh | ((uint32_t)hit << 16)) : 0u);
        const uint32_t f = fk & 0xffffu, hk = fk >> 16, tot = hdr & 0x00ffffffu;
        rc.encode(sumf + hl * SFQ_QS + hk, f + 1u, tot + 64u);
        update(hl, hk, f, tot, hdr >> 24);
    }

    // Log64Ranger::get (log64_ranger.hpp:114-138) on the loaded context.
    __device__ __forceinline__ uint32_t get(SfqDec &rc) {
        const uint32_t hdr = header();
        const uint32_t tot = hdr & 0x00ffffffu;
        const uint32_t prob = rc.get_freq(tot + 64u);
        uint32_t lsum = 0;
#pragma unroll
        for (int k = 0; k < SFQ_QS; k++) lsum += (w[k] & 0xffffu) + 1u;
        uint32_t excl = 0;                                   // 7 independent shuffles instead of a 3-step dependent scan
#pragma unroll
        for (int d = 1; d < SFQ_QG; d++) { const uint32_t t = __shfl_up_sync(gmask, lsum, d, SFQ_QG); if ((int)lane >= d) excl += t; }
        const unsigned ball = (__ballot_sync(gmask, prob < excl + lsum) >> gbase) & 0xffu;
        const uint32_t hl = ball ? (uint32_t)(__ffs(ball) - 1) : (uint32_t)(SFQ_QG - 1);     // no lane: corrupt stream
        uint32_t cum = excl, fh = 0, sh = 0;
        int hit = -1;
#pragma unroll
        for (int k = 0; k < SFQ_QS; k++) {
            const uint32_t f = w[k] & 0xffffu;
            if (hit < 0) {
                if (cum + f + 1u <= prob && k < SFQ_QS - 1) cum += f + 1u;
                else { hit = k; fh = f; sh = ((w[k] >> 16) & 0xffu) ^ (lane * SFQ_QS + k); }
            }
        }
        const uint32_t sumf = __reduce_or_sync(gmask, lane == hl ? cum : 0u);
        const uint32_t pk = __reduce_or_sync(gmask, lane == hl ? (fh | ((uint32_t)hit << 16) | (sh << 19)) : 0u);
        const uint32_t f = pk & 0xffffu, hk = (pk >> 16) & 7u, sym = pk >> 19;
        rc.decode(sumf, f + 1u);
        update(hl, hk, f, tot, hdr >> 24);
        return sym;
    }
};

// One group of SFQ_QG lanes per chunk.
__device__ __forceinline__ void sfq_qlt_encode_group(const uint8_t *text, const uint64_t *ls, SfqChunkMeta *meta, int level,
                                                     uint32_t *qtable, uint32_t nent, uint32_t *pwpool, uint8_t *arena,
                                                     SfqArena *ar, SfqQGroup &g) {
    SfqEnc rc;
    rc.start(arena + ar->off[SFQ_S_QLT], ar->cap[SFQ_S_QLT]);
    rc.out.mute = (g.lane != 0);
    SfqPower ex; ex.m = pwpool + (size_t)SFQ_PW_QEX * SFQ_PW_WORDS;
    SfqQTable tab;
    tab.init(qtable, nent, level <= 1 || nent >= 65536u);
    const uint32_t solid = meta->solid;
    uint32_t extra_hi = 0;
    bool full = false;
    SfqQltCursor ahead;                      // run-ahead prefetch: every lane touches its own sector
    ahead.init(text, ls, meta, level);
    ahead.run(tab, g.lane, SFQ_QLT_AHEAD);
    for (uint32_t r = 0; r < meta->nrec && !full; r++) {
        const SfqRecView v = sfq_rec_view(text, ls, meta->line0, r, solid);
        SfqQCtx c; c.reset();
        SfqReader rq;
        if (v.qlen) rq.seek(v.qual);
        for (uint32_t i = 0; i < v.qlen; i++) {
            if ((i & 15u) == 0) ahead.run(tab, g.lane, 16);
            const uint8_t b = (uint8_t)(rq.next() - '!');
            if (!g.locate(tab, c.last)) { full = true; break; }
            if (b < 63) g.put(rc, b);
            else {                                                              // qlts.cpp:120-125
                g.put(rc, 63);
                if (g.lane == 0) ex.put(rc, b);            // one lane runs the serial 256-symbol model ...
                __syncwarp(g.gmask);
                rc.low = ((uint64_t)g.bcast((uint32_t)(rc.low >> 32), 0) << 32) | g.bcast((uint32_t)rc.low, 0);
                rc.range = g.bcast(rc.range, 0);           // ... and hands the coder state back to the group
                rc.out.n = g.bcast(rc.out.n, 0);
                extra_hi++;
            }
            sfq_q_next(c, level, b);
        }
    }
    rc.finish();
    if (g.lane == 0) {
        ar->size[SFQ_S_QLT] = rc.out.n;
        meta->extra_hi = extra_hi;
        meta->q_used = tab.used;
        if (meta->status == SFQ_OK) { if (full) meta->status = SFQ_E_TABLE; else if (rc.out.overflow()) meta->status = SFQ_E_CAP; }
    }
}

__device__ __forceinline__ void sfq_qlt_decode_group(const uint8_t *in, const uint32_t *ssize, const uint64_t *soff,
                                                     SfqChunkMeta *meta, int level, uint32_t *qtable, uint32_t nent, uint32_t *pwpool,
                                                     const uint32_t *qlen_tab, const uint64_t *qoff_tab, uint8_t *quals, SfqQGroup &g) {
    SfqDec rc;
    rc.start(in + soff[SFQ_S_QLT], ssize[SFQ_S_QLT]);
    SfqPower ex; ex.m = pwpool + (size_t)SFQ_PW_QEX * SFQ_PW_WORDS;
    SfqQTable tab;
    tab.init(qtable, nent, level <= 1 || nent >= 65536u);
    bool full = false;
    for (uint32_t r = 0; r < meta->nrec && !full; r++) {
        const uint32_t qlen = sfq_coded_len(qlen_tab[r]);
        uint8_t *q = quals + qoff_tab[r];
        SfqQCtx c; c.reset();
        for (uint32_t i = 0; i < qlen; i++) {
            if (!g.locate(tab, c.last)) { full = true; break; }
            uint32_t b = g.get(rc);
            if (b == 63) {                                                      // qlts.cpp:206-208
                if (g.lane == 0) b = ex.get(rc);
                __syncwarp(g.gmask);
                b = g.bcast(b, 0);
                rc.low = ((uint64_t)g.bcast((uint32_t)(rc.low >> 32), 0) << 32) | g.bcast((uint32_t)rc.low, 0);
                rc.code = ((uint64_t)g.bcast((uint32_t)(rc.code >> 32), 0) << 32) | g.bcast((uint32_t)rc.code, 0);
                rc.range = g.bcast(rc.range, 0);
                rc.pos = g.bcast(rc.pos, 0);
                rc.have = false;
            }
            if (g.lane == 0) q[i] = (uint8_t)('!' + b);
            sfq_q_next(c, level, (uint8_t)b);
        }
    }
    if (full && g.lane == 0 && meta->status == SFQ_OK) meta->status = SFQ_E_TABLE;
}

#endif  // __CUDACC__
