// Quality stream decoder, warp-converged form: four chunks per warp, eight lanes ("an octet") per chunk.
//
// QltLoad::load_1/2/3 (qlts.cpp:163-234) is a strict serial chain: the context of symbol n+1 depends on
// symbol n, so one chunk-stream cannot be split.  What can be cut is the time of one link of the chain.
// The first cooperative decoder (sfq_qlt_group.cuh) gave every octet its own sub-warp mask; the compiler
// then guards each shuffle with a WARPSYNC/collective sequence and the four octets of a warp run one after
// another.  Here all 32 lanes stay converged for the whole chunk: every warp primitive uses the full mask
// (width-8 segments), every branch is decided by a warp vote, rare cases are predicated per octet.
//
// Model of one context (Log64Ranger, log64_ranger.hpp:36-140) = 256 bytes = 8 lanes x 32 bytes; lane L holds
//   words 0..3   freq[8L .. 8L+7]            (u16 each)
//   words 4..5   symbol of slot 8L+k XOR (8L+k), one byte each  -> zeroed memory is the reference's start state
//   word  6      total (bits 0..21) | count (bits 24..31), replicated in every lane
//   word  7      sum of freq over the slots of lanes < L; in lane 0 (whose sum is 0 by definition) the hash key.
//                An entry is occupied once its total is non-zero - every update adds to it
// so a lane needs nothing from its neighbours to place `code` among its eight cumulative frequencies: one
// division (range / total), eight multiply-compares, one ballot to pick the lane, two shuffles to broadcast
// the slot.  range / totFreq and the search "largest cum with cum * r <= code" replace the reference's
// 64-bit code / range (coder.hpp:83-86): floor(code / r) >= c  <=>  code >= c * r.
#pragma once
#include "sfq_streams.cuh"

#if defined(__CUDACC__)

// n = 1..3 bytes of the stream, first byte most significant
__device__ __forceinline__ uint32_t sfq_src_take(SfqByteSrc &s, uint32_t n) {
    uint32_t v;
    if (s.left >= n) { v = (uint32_t)s.word; s.word >>= 8u * n; s.left -= n; }
    else {
        const uint64_t nw = s.ahead;
        s.ahead = s.fetch();
        v = (uint32_t)s.word | (uint32_t)(nw << (8u * s.left));
        const uint32_t rest = n - s.left;
        s.word = nw >> (8u * rest);
        s.left = 8u - rest;
    }
    return __byte_perm(v, 0u, 0x0123u) >> (8u * (4u - n));
}

// Coder state of one chunk-stream (replicated in the eight lanes of its octet).  get_freq / decode are the
// reference-shaped forms used by the rare paths (escape symbols through the 256-symbol model).
struct SfqQdCoder {
    uint64_t low, code;
    uint32_t range;
    SfqByteSrc src;
    __device__ __forceinline__ void idle() { low = 0; code = 0; range = 0xFFFFFFFFu; src.p = nullptr; src.end = nullptr; src.word = 0; src.ahead = 0; src.left = 8; }
    __device__ __forceinline__ void start(const uint8_t *buf, uint32_t size) {
        low = 0; code = 0; range = 0xFFFFFFFFu;
        src.start(buf, size);
        for (int k = 0; k < 8; k++) code = (code << 8) | src.next();               // coder.hpp:44-48
    }
    __device__ __forceinline__ uint32_t get_freq(uint32_t tot) {
        range /= tot;
        return (code >> 32) ? (uint32_t)(code / range) : ((uint32_t)code / range);
    }
    __device__ __forceinline__ void renorm_exact() {                               // coder.hpp:92-101
        while (range < SFQ_RC_TOP) {
            if ((low ^ (low + range)) & (0xffULL << 56))
                range = (((uint32_t)low) | (SFQ_RC_TOP - 1)) - (uint32_t)low;
            code = (code << 8) | src.next();
            range <<= 8;
            low <<= 8;
        }
    }
    __device__ __forceinline__ void decode(uint32_t cum, uint32_t freq) {
        const uint32_t t = cum * range;
        low += t; code -= t; range *= freq;
        renorm_exact();
    }
};

// Decoded qualities of a chunk are contiguous in the plane: eight bytes are gathered in a register and
// stored as one aligned word (the chunk's first and last partial words go out bytewise).
struct SfqQdSink {
    uint8_t *p;            // next byte
    uint64_t acc;          // bytes of the aligned word under p written so far
    __device__ __forceinline__ void start(uint8_t *first) { p = first; acc = 0; }
};

struct SfqQdModel { uint32_t f0, f1, f2, f3, f4, f5, f6, f7, sx, sy, hdr, excl; };
__device__ __forceinline__ void sfq_qd_load(SfqQdModel &m, const uint32_t *e) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(e)), b = __ldcg(reinterpret_cast<const uint4 *>(e + 4));
    m.f0 = a.x & 0xffffu; m.f1 = a.x >> 16; m.f2 = a.y & 0xffffu; m.f3 = a.y >> 16;
    m.f4 = a.z & 0xffffu; m.f5 = a.z >> 16; m.f6 = a.w & 0xffffu; m.f7 = a.w >> 16;
    m.sx = b.x; m.sy = b.y; m.hdr = b.z; m.excl = b.w;
}
__device__ __forceinline__ void sfq_qd_store_freqs(const SfqQdModel &m, uint32_t *e) {
    *reinterpret_cast<uint4 *>(e) = make_uint4(m.f0 | (m.f1 << 16), m.f2 | (m.f3 << 16), m.f4 | (m.f5 << 16), m.f6 | (m.f7 << 16));
}
__device__ __forceinline__ void sfq_qd_store_tail(const SfqQdModel &m, uint32_t *e) {
    *reinterpret_cast<uint4 *>(e + 4) = make_uint4(m.sx, m.sy, m.hdr, m.excl);
}
__device__ __forceinline__ void sfq_qd_store_hdr(const SfqQdModel &m, uint32_t *e) {
    *reinterpret_cast<uint2 *>(e + 6) = make_uint2(m.hdr, m.excl);
}

#define SFQ_QD_WARPS 2                  // warps per CTA
__global__ void __launch_bounds__(32 * SFQ_QD_WARPS)
k_qlt_decode4(const uint8_t *__restrict__ in, const SfqDecChunk *__restrict__ dc, SfqChunkMeta *metas, SfqWorkspace ws,
              SfqRecTables t, uint8_t *quals, uint32_t nchunks) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, l8 = lane & 7u, osh = lane & 24u;
    const uint32_t c = (blockIdx.x * SFQ_QD_WARPS + (threadIdx.x >> 5)) * 4u + (lane >> 3);
    bool live = c < nchunks;
    if (live) live = metas[c].status == SFQ_OK;
    const bool valid = live;

    uint32_t nrec = 0, nent = 1, level = 3;
    bool dense = true;
    uint32_t *tab = ws.qtab, *pwq = ws.pw;
    const uint32_t *qlen_tab = t.qlen;
    uint8_t *oplane = quals;
    SfqQdCoder rc;
    rc.idle();
    if (live) {
        const SfqDecChunk &d = dc[c];
        nrec = metas[c].nrec;
        level = (uint32_t)d.level;
        nent = ws.cbits;
        dense = level <= 1u || nent >= 65536u;
        tab = ws.qtab + (size_t)c * ws.qtab_words;
        pwq = ws.pw + ((size_t)c * SFQ_PW_PER_CHUNK + SFQ_PW_QEX) * SFQ_PW_WORDS;
        qlen_tab = t.qlen + d.rec_base;
        rc.start(in + d.soff[SFQ_S_QLT], d.ssize[SFQ_S_QLT]);
        oplane = quals + d.qual_plane;
    }
    tab += l8 * 8u;                                             // this lane's 32 bytes of every entry
    // decoded qualities of a chunk are contiguous in the plane: eight bytes are gathered in two registers and
    // stored as one aligned word; `opos` counts from the aligned address at or below the chunk's first byte
    uint8_t *const obase = (uint8_t *)((uintptr_t)oplane & ~(uintptr_t)7);
    const uint32_t ohead = (uint32_t)((uintptr_t)oplane & 7u);
    uint32_t opos = ohead, alo = 0, ahi = 0;
    const uint32_t lmask = level <= 1u ? 0xfffu : 0xffffu;

    // record cursor: r = record, i = position inside it; empty records are skipped
    uint32_t r = 0, i = 0, qlen = 0;
#define SFQ_QD_OPEN(want)                                                                   \
    {                                                                                        \
        bool w_ = (want);                                                                    \
        while (__any_sync(FULL, w_)) {                                                       \
            if (w_) {                                                                        \
                if (r >= nrec) { live = false; w_ = false; }                                 \
                else { qlen = qlen_tab[r]; if (qlen) { i = 0; w_ = false; } else r++; }       \
            }                                                                                \
        }                                                                                    \
    }
    SFQ_QD_OPEN(live)

    // context state (qlts.hpp:52-74, qlts.cpp:109-134): `last` for levels 1-2; previous / one-before symbols
    // and the running drop sum for levels 3-4
    uint32_t last = 0, p1 = 0, p2 = 0, dl = 5;
    uint32_t ctx = 0, h = 0, used = 0;
    bool full = false;
    SfqQdModel m;
    m.f0 = m.f1 = m.f2 = m.f3 = m.f4 = m.f5 = m.f6 = m.f7 = 0; m.sx = m.sy = m.hdr = m.excl = 0;
    uint4 na = make_uint4(0, 0, 0, 0), nb = make_uint4(0, 0, 0, 0);      // the next model, as loaded
    if (live) { na = __ldcg(reinterpret_cast<const uint4 *>(tab)); nb = __ldcg(reinterpret_cast<const uint4 *>(tab + 4)); }
    bool fresh = live;                                                    // na/nb hold a model not yet unpacked
    bool chk = live && !dense;

    while (__any_sync(FULL, live)) {
        if (fresh) {
            m.f0 = na.x & 0xffffu; m.f1 = na.x >> 16; m.f2 = na.y & 0xffffu; m.f3 = na.y >> 16;
            m.f4 = na.z & 0xffffu; m.f5 = na.z >> 16; m.f6 = na.w & 0xffffu; m.f7 = na.w >> 16;
            m.sx = nb.x; m.sy = nb.y; m.hdr = nb.z; m.excl = nb.w;
        }
        // ---------------------------------------------------------------- the entry of `ctx` (hash probe)
        // An entry is in use once its total is non-zero (every update adds to it, and every lane holds the
        // total); lane 0 keeps the 16-bit key in its otherwise unused prefix word.
        if (__any_sync(FULL, chk)) {
            uint32_t probes = 0;
            for (;;) {
                const bool occ = (m.hdr & 0x3fffffu) != 0u;
                const unsigned vb = __ballot_sync(FULL, chk && occ && l8 == 0 && m.excl != ctx);
                const bool bad = (vb >> osh) & 1u;
                if (chk && !occ) {                                        // first visit of this context in the chunk
                    if (used + 1u >= nent) { full = true; live = false; }
                    else { used++; if (l8 == 0) m.excl = ctx; }
                    chk = false;
                } else if (chk && bad) {                                  // somebody else's entry: walk on
                    h = h + 1u == nent ? 0u : h + 1u;
                    if (++probes > nent) { full = true; live = false; chk = false; }
                    else sfq_qd_load(m, tab + (size_t)h * 64u);
                } else chk = false;
                if (!vb) break;
            }
        }
        const bool act = live;

        // ---------------------------------------------------------------- Log64Ranger::get (log64_ranger.hpp:114-138)
        const uint32_t tot = m.hdr & 0x3fffffu, count = m.hdr >> 24;
        const uint32_t rr = rc.range / (tot + 64u);
        const uint32_t e0 = (l8 ? m.excl : 0u) + 8u * l8;
        const uint32_t c0 = e0 + m.f0 + 1u, c1 = c0 + m.f1 + 1u, c2 = c1 + m.f2 + 1u, c3 = c2 + m.f3 + 1u,
                       c4 = c3 + m.f4 + 1u, c5 = c4 + m.f5 + 1u, c6 = c5 + m.f6 + 1u, c7 = c6 + m.f7 + 1u;
        // (a corrupt stream can leave code >= 2^32: every comparison below is then true, as with the reference's 64-bit quotient)
        const uint32_t code32 = (uint32_t)(rc.code >> 32) ? 0xffffffffu : (uint32_t)rc.code;
        const unsigned sel = (__ballot_sync(FULL, code32 < c7 * rr) >> osh) & 0xffu;
        const uint32_t hl = sel ? (uint32_t)(__ffs(sel) - 1) : 7u;          // no lane: corrupt stream, last slot
        uint32_t cb, hk, pack;
        {   // binary search of code among this lane's thresholds c0..c6 (meaningful in lane hl)
            const bool g3 = code32 >= c3 * rr;
            const uint32_t cm = g3 ? c5 : c1;
            uint32_t lo = g3 ? c3 : e0, hi = g3 ? c7 : c3;
            const bool g2 = code32 >= cm * rr;
            const uint32_t cq = g3 ? (g2 ? c6 : c4) : (g2 ? c2 : c0);
            lo = g2 ? cm : lo; hi = g2 ? hi : cm;
            const bool g1 = code32 >= cq * rr;
            lo = g1 ? cq : lo; hi = g1 ? hi : cq;
            hk = (g3 ? 4u : 0u) + (g2 ? 2u : 0u) + (g1 ? 1u : 0u);
            const uint32_t sb = __byte_perm(m.sx, m.sy, hk) & 0xffu;
            cb = lo;
            pack = (hi - lo - 1u) | ((sb ^ (8u * l8 + hk)) << 16) | (hk << 24);
        }
        cb = __shfl_sync(FULL, cb, hl, 8);                                   // the chosen slot, from its lane
        pack = __shfl_sync(FULL, pack, hl, 8);
        const uint32_t f = pack & 0xffffu, ssym = (pack >> 16) & 0xffu;
        hk = pack >> 24;
        {   // Decode (coder.hpp:88-91)
            const uint32_t tt = cb * rr;
            rc.low += tt;
            rc.code -= tt;
            rc.range = rr * (f + 1u);
        }
        {   // renormalise: n whole bytes at once.  The carry guard (coder.hpp:95-96) can only fire when bits
            // 32..39 of low are all ones (for any of the n byte steps); then take the reference's loop.
            const uint32_t n = act ? (uint32_t)__clz((int)rc.range) >> 3 : 0u;
            const bool risky = n != 0u && ((uint32_t)(rc.low >> 32) & 0xffu) == 0xffu;
            if (n != 0u && !risky) {
                const uint32_t v = sfq_src_take(rc.src, n);
                rc.code = (rc.code << (8u * n)) | v;
                rc.range <<= 8u * n;
                rc.low <<= 8u * n;
            }
            if (__any_sync(FULL, risky)) { if (risky) rc.renorm_exact(); }
        }
        uint32_t b = ssym;

        // ---------------------------------------------------------------- escape: 63 + 256-symbol model (qlts.cpp:206-208)
        if (__any_sync(FULL, act && b == 63u)) {
            const bool esc = act && b == 63u;
            if (esc && l8 == 0) { SfqPower ex; ex.m = pwq; b = ex.get(rc); }
            __syncwarp();
#define SFQ_QD_BC32(x) { const uint32_t v_ = __shfl_sync(FULL, (uint32_t)(x), 0, 8); if (esc) x = v_; }
#define SFQ_QD_BC64(x) { const uint32_t lo_ = __shfl_sync(FULL, (uint32_t)(x), 0, 8), hi_ = __shfl_sync(FULL, (uint32_t)((uint64_t)(x) >> 32), 0, 8); \
                         if (esc) x = ((uint64_t)hi_ << 32) | lo_; }
            uint64_t sp = (uint64_t)(uintptr_t)rc.src.p;
            SFQ_QD_BC32(b) SFQ_QD_BC32(rc.range) SFQ_QD_BC32(rc.src.left)
            SFQ_QD_BC64(rc.low) SFQ_QD_BC64(rc.code) SFQ_QD_BC64(rc.src.word) SFQ_QD_BC64(rc.src.ahead) SFQ_QD_BC64(sp)
            rc.src.p = (const uint8_t *)(uintptr_t)sp;
        }

        // ---------------------------------------------------------------- output, next position, next context
        if (act) {
            const uint32_t k = opos & 7u;
            const uint32_t v = ((b + 33u) & 0xffu) << (8u * (k & 3u));
            if (k & 4u) ahi |= v; else alo |= v;
            opos++;
            if (k == 7u) {
                if (l8 == 0) {
                    if (opos - 8u >= ohead) *reinterpret_cast<uint2 *>(obase + (opos - 8u)) = make_uint2(alo, ahi);
                    else for (uint32_t q = ohead; q < 8u; q++) obase[q] = (uint8_t)((q & 4u ? ahi : alo) >> (8u * (q & 3u)));
                }
                alo = 0; ahi = 0;
            }
            i++;
        }
        const bool endrec = act && i == qlen;
        {   // qlts.hpp:52-74; b is not clamped when it feeds the context
            const uint32_t b8 = b & 0xffu;
            const uint32_t l12 = (b8 | (last << 6)) & lmask;
            dl += p1 > b8 ? p1 - b8 : 0u;
            const uint32_t d3 = dl >> 3;
            const uint32_t l3 = (b8 | ((p1 > p2 ? p1 : p2) << 6) | ((p1 == p2 ? 1u : 0u) << 12) | ((d3 < 7u ? d3 : 7u) << 13)) & 0xffffu;
            p2 = p1; p1 = b8;
            last = level >= 3u ? l3 : l12;
            if (endrec) { last = 0; p1 = 0; p2 = 0; dl = 5; r++; }
        }
        SFQ_QD_OPEN(endrec)
        const uint32_t nctx = last;

        // ---------------------------------------------------------------- request the next model before updating this one
        fresh = live && nctx != ctx;
        uint32_t hn = h;
        if (fresh) {
            hn = dense ? nctx : __umulhi(nctx * 2654435761u, nent);
            if (!dense && hn == h) hn = h + 1u == nent ? 0u : h + 1u;       // entry h is ours (key = ctx): skip it unseen, its store is still pending
            const uint32_t *e = tab + (size_t)hn * 64u;
            na = __ldcg(reinterpret_cast<const uint4 *>(e));
            nb = __ldcg(reinterpret_cast<const uint4 *>(e + 4));
        }

        // ---------------------------------------------------------------- update_freq (log64_ranger.hpp:69-87)
        const uint32_t slot = 8u * hl + hk;
        bool skip = false, wide = false;
        uint32_t fn = f, tot2 = tot;
        if (__any_sync(FULL, act && f > 65472u - 6u)) {
            const bool hv = act && f > 65472u - 6u;
            if (hv && slot == 0u && f + 20u > tot) skip = true;               // saturated front slot: no update at all
            const bool dohalve = hv && !skip;
            if (dohalve) { m.f0 >>= 1; m.f1 >>= 1; m.f2 >>= 1; m.f3 >>= 1; m.f4 >>= 1; m.f5 >>= 1; m.f6 >>= 1; m.f7 >>= 1; }
            const uint32_t ls = m.f0 + m.f1 + m.f2 + m.f3 + m.f4 + m.f5 + m.f6 + m.f7;
            uint32_t inc = ls, tmp;
            tmp = __shfl_up_sync(FULL, inc, 1, 8); if (l8 >= 1u) inc += tmp;
            tmp = __shfl_up_sync(FULL, inc, 2, 8); if (l8 >= 2u) inc += tmp;
            tmp = __shfl_up_sync(FULL, inc, 4, 8); if (l8 >= 4u) inc += tmp;
            const uint32_t total = __shfl_sync(FULL, inc, 7, 8);
            if (dohalve) {
                if (l8) m.excl = inc - ls;
                tot2 = total;
                fn = f >> 1;
                wide = true;
            }
        }
        const bool upd = act && !skip;
        const bool mine = upd && l8 == hl;
        if (upd) {
            fn += 6u;
            tot2 += 6u;
            if (l8 > hl) m.excl += 6u;
        }
        if (mine) {
            m.f0 = hk == 0u ? fn : m.f0; m.f1 = hk == 1u ? fn : m.f1; m.f2 = hk == 2u ? fn : m.f2; m.f3 = hk == 3u ? fn : m.f3;
            m.f4 = hk == 4u ? fn : m.f4; m.f5 = hk == 5u ? fn : m.f5; m.f6 = hk == 6u ? fn : m.f6; m.f7 = hk == 7u ? fn : m.f7;
        }
        uint32_t cnt2 = count;
        const bool cand = upd && slot != 0u;                                 // `++count` is not evaluated for slot 0
        if (cand) cnt2 = (count + 1u) & 0xffu;
        const bool swapc = cand && (cnt2 & 0xfu) == 0u;
        if (__any_sync(FULL, swapc)) {                                       // maybe swap slot with slot-1 (down_level)
            const uint32_t lf7 = __shfl_up_sync(FULL, m.f7, 1, 8);            // left neighbour's last slot
            const uint32_t lb7 = __shfl_up_sync(FULL, m.sy >> 24, 1, 8);
            uint64_t s64 = ((uint64_t)m.sy << 32) | m.sx;
            // lane hl works out the neighbour slot and the verdict, the octet hears it
            uint32_t fprev = 0, bprev = 0;
            if (hk == 0u) { fprev = lf7; bprev = lb7; }
            else {
                fprev = hk == 1u ? m.f0 : hk == 2u ? m.f1 : hk == 3u ? m.f2 : hk == 4u ? m.f3 : hk == 5u ? m.f4 : hk == 6u ? m.f5 : m.f6;
                bprev = (uint32_t)(s64 >> (8u * (hk - 1u))) & 0xffu;
            }
            uint32_t verdict = (swapc && fn > fprev ? 1u : 0u) | (fprev << 1) | (bprev << 24);
            verdict = __shfl_sync(FULL, verdict, hl, 8);
            if (verdict & 1u) {
                fprev = (verdict >> 1) & 0xffffu; bprev = verdict >> 24;
                const uint32_t symprev = bprev ^ (slot - 1u);
                const uint32_t byte_lo = (ssym ^ (slot - 1u)) & 0xffu;        // slot-1 now holds this symbol ...
                const uint32_t byte_hi = (symprev ^ slot) & 0xffu;            // ... and slot the neighbour's
                if (l8 == hl) {
                    if (hk == 0u) {
                        m.f0 = fprev;
                        s64 = (s64 & ~0xffull) | byte_hi;
                        m.excl += fn - fprev;                                 // (hl >= 1 here: lane 0's word keeps the key)
                    } else {
                        const uint32_t k1 = hk - 1u;
                        m.f0 = k1 == 0u ? fn : m.f0;                            m.f1 = k1 == 1u ? fn : hk == 1u ? fprev : m.f1;
                        m.f2 = k1 == 2u ? fn : hk == 2u ? fprev : m.f2; m.f3 = k1 == 3u ? fn : hk == 3u ? fprev : m.f3;
                        m.f4 = k1 == 4u ? fn : hk == 4u ? fprev : m.f4; m.f5 = k1 == 5u ? fn : hk == 5u ? fprev : m.f5;
                        m.f6 = k1 == 6u ? fn : hk == 6u ? fprev : m.f6; m.f7 = hk == 7u ? fprev : m.f7;
                        s64 = (s64 & ~(0xffffull << (8u * k1))) | ((uint64_t)(byte_lo | (byte_hi << 8)) << (8u * k1));
                    }
                    wide = true;
                }
                if (hk == 0u && l8 + 1u == hl) {
                    m.f7 = fn;
                    s64 = (s64 & ~(0xffull << 56)) | ((uint64_t)byte_lo << 56);
                    wide = true;
                }
                m.sx = (uint32_t)s64; m.sy = (uint32_t)(s64 >> 32);
            }
        }
        if (act) {
            m.hdr = tot2 | (cnt2 << 24);
            uint32_t *e = tab + (size_t)h * 64u;
            if (wide || mine) sfq_qd_store_freqs(m, e);
            if (wide) sfq_qd_store_tail(m, e); else sfq_qd_store_hdr(m, e);
        }
        if (fresh) { h = hn; ctx = nctx; chk = !dense; }
    }
#undef SFQ_QD_OPEN
#undef SFQ_QD_BC32
#undef SFQ_QD_BC64
    if (valid && l8 == 0) {
        // the last partial word
        const uint32_t w0 = opos & ~7u;
        for (uint32_t q = w0 < ohead ? ohead : w0; q < opos; q++) obase[q] = (uint8_t)((q & 4u ? ahi : alo) >> (8u * (q & 3u)));
        if (full && metas[c].status == SFQ_OK) metas[c].status = SFQ_E_TABLE;
    }
}

#endif  // __CUDACC__
