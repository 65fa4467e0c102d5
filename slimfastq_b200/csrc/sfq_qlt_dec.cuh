// Quality stream decoder, warp-converged form: LPC lanes per chunk (eight - four chunks per warp - or four).
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

// One context's model as a lane holds it.  LPC = lanes per chunk (8 or 4), SPL = 64 / LPC slots per lane.
// In memory a lane owns WPL = 8 (LPC 8) or 16 (LPC 4) consecutive words of the 64-word entry:
//   SPL/2 words of freq pairs | SPL/4 words of symbol bytes | hdr | excl | (padding)
template <int LPC> struct SfqQdGeo {
    static constexpr int SPL = 64 / LPC;                    // slots per lane
    static constexpr int WPL = 64 / LPC;                    // words per lane (8 or 16)
    static constexpr int FW = SPL / 2, SW = SPL / 4;        // words of frequencies / of symbol bytes
    static constexpr int HW = FW + SW;                      // word index of hdr (excl follows)
    static constexpr int NV = WPL / 4;                      // 16-byte vectors per lane
    static constexpr int CPW = 32 / LPC;                    // chunks per warp
};
template <int LPC> struct SfqQdModelT {
    typedef SfqQdGeo<LPC> G;
    uint32_t pw[G::FW];                 // the frequencies as stored (u16 pairs): the state carried from step to step
    uint32_t f[G::SPL];                 // ... and taken apart: rebuilt from pw at the top of every step (derive), so the common
                                        // update is one add into a packed word instead of a 16-way select + repacking
    uint32_t sy[G::SW];
    uint32_t hdr, excl;
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < G::FW; k++) pw[k] = 0;
#pragma unroll
        for (int k = 0; k < G::SPL; k++) f[k] = 0;
#pragma unroll
        for (int k = 0; k < G::SW; k++) sy[k] = 0;
        hdr = 0; excl = 0;
    }
    __device__ __forceinline__ void derive() {
#pragma unroll
        for (int k = 0; k < G::FW; k++) { f[2 * k] = pw[k] & 0xffffu; f[2 * k + 1] = pw[k] >> 16; }
    }
    __device__ __forceinline__ void pack() {
#pragma unroll
        for (int k = 0; k < G::FW; k++) pw[k] = f[2 * k] | (f[2 * k + 1] << 16);
    }
    // freq[k] += inc in the packed words (the caller knows the sum stays below 2^16)
    __device__ __forceinline__ void add_packed(uint32_t k, uint32_t inc) {
        const uint32_t v = inc << (16u * (k & 1u));
#pragma unroll
        for (int q = 0; q < G::FW; q++) pw[q] += (k >> 1) == (uint32_t)q ? v : 0u;
    }
    // word j (a compile-time index after unrolling) of the lane's vectors: stays in registers
    static __device__ __forceinline__ uint32_t word(const uint4 (&v)[G::NV], int j) {
        const uint4 &q = v[j >> 2];
        return (j & 3) == 0 ? q.x : (j & 3) == 1 ? q.y : (j & 3) == 2 ? q.z : q.w;
    }
    __device__ __forceinline__ void unpack(const uint4 (&v)[G::NV]) {      // (pw only: derive() follows before f is used)
#pragma unroll
        for (int k = 0; k < G::FW; k++) pw[k] = word(v, k);
#pragma unroll
        for (int k = 0; k < G::SW; k++) sy[k] = word(v, G::FW + k);
        hdr = word(v, G::HW); excl = word(v, G::HW + 1);
    }
    __device__ __forceinline__ void store_freqs(uint32_t *e) const {
#pragma unroll
        for (int q = 0; q < G::FW / 4; q++)
            *reinterpret_cast<uint4 *>(e + 4 * q) = make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
    }
    __device__ __forceinline__ void store_syms(uint32_t *e) const {
        if (G::SW == 2) *reinterpret_cast<uint2 *>(e + G::FW) = make_uint2(sy[0], sy[1]);
        else *reinterpret_cast<uint4 *>(e + G::FW) = make_uint4(sy[0], sy[G::SW > 1 ? 1 : 0], sy[G::SW > 2 ? 2 : 0], sy[G::SW > 3 ? 3 : 0]);
    }
    __device__ __forceinline__ void store_hdr(uint32_t *e) const { *reinterpret_cast<uint2 *>(e + G::HW) = make_uint2(hdr, excl); }
    // Runtime-indexed access to the register arrays, written as mask arithmetic over every element: a chain of
    // selects or conditional stores gets turned into an indexed access by the compiler, which would move the
    // whole model to local memory.
    __device__ __forceinline__ uint32_t sym_byte(uint32_t k) const {          // stored byte of slot k of this lane
        uint32_t w = 0;
#pragma unroll
        for (int q = 0; q < G::SW; q++) w |= sy[q] & (0u - (uint32_t)((k >> 2) == (uint32_t)q));
        return (w >> (8u * (k & 3u))) & 0xffu;
    }
    __device__ __forceinline__ void set_sym_byte(uint32_t k, uint32_t v) {
        const uint32_t sh = 8u * (k & 3u);
#pragma unroll
        for (int q = 0; q < G::SW; q++) {
            const uint32_t hit = 0u - (uint32_t)((k >> 2) == (uint32_t)q);
            sy[q] = (sy[q] & ~((0xffu << sh) & hit)) | ((v << sh) & hit);
        }
    }
    __device__ __forceinline__ uint32_t freq_at(uint32_t k) const {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < G::SPL; j++) v |= f[j] & (0u - (uint32_t)(k == (uint32_t)j));
        return v;
    }
    __device__ __forceinline__ void set_freq(uint32_t k, uint32_t v) {
#pragma unroll
        for (int j = 0; j < G::SPL; j++) { const uint32_t hit = 0u - (uint32_t)(k == (uint32_t)j); f[j] = (f[j] & ~hit) | (v & hit); }
    }
};
template <int LPC>
__device__ __forceinline__ void sfq_qd_fetch(uint4 (&v)[SfqQdGeo<LPC>::NV], const uint32_t *e) {
#pragma unroll
    for (int q = 0; q < SfqQdGeo<LPC>::NV; q++) v[q] = __ldcg(reinterpret_cast<const uint4 *>(e + 4 * q));
}

// CH ("compact header", LPC 4 only): everything a visit changes besides one lane's frequencies lives in ONE 32-byte sector
// of the entry instead of in every lane's own words -
//   words 0..7    hdr (total | count << 24), prefix of lanes 1, 2, 3, key, 3 unused
//   words 8..39   freq pairs, lane L at 8 + 8L           words 40..55  symbol bytes, lane L at 40 + 4L
// - so a visit reads 7 sectors and dirties 2 (the header's and the coded slot's lane) where the lane-owned layout reads 8 and
// dirties 5.  Measured (profiles/README.md, r2v): the kernel's DRAM traffic falls by 30 %, its link grows by 9 % (eight load
// instructions instead of four, the lane's prefix picked by selects) and at the benched residency it ends up 3 % slower -
// bit-exact, kept for A/B (SFQ_QCH=1), not the default.
#define SFQ_QD_MAXW 8                   // most warps per CTA (the launch picks 2 or 8)
template <int LPC, bool SPEC, bool CH = false>
__global__ void __launch_bounds__(32 * SFQ_QD_MAXW)
k_qlt_decode(const uint8_t *__restrict__ in, const SfqDecChunk *__restrict__ dc, SfqChunkMeta *metas, SfqWorkspace ws,
             SfqRecTables t, uint8_t *quals, uint32_t nchunks) {
    typedef SfqQdGeo<LPC> G;
    constexpr uint32_t SPL = G::SPL;
    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, l8 = lane % LPC, osh = lane - l8;       // l8 = lane inside the chunk's group
    const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * G::CPW + lane / LPC;
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
        oplane = quals + d.qual_plane + metas[c].big_quals;      // (oversized records' quality lines sit at the front of the chunk's plane)
    }
    static_assert(!CH || (LPC == 4 && !SPEC), "compact-header layout: 4 lanes per chunk, no speculative prefetch");
    if (!CH) tab += l8 * G::WPL;                                 // this lane's words of every entry
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
                else { qlen = sfq_coded_len(qlen_tab[r]); if (qlen) { i = 0; w_ = false; } else r++; } \
            }                                                                                \
        }                                                                                    \
    }
    SFQ_QD_OPEN(live)

    // context state (qlts.hpp:52-74, qlts.cpp:109-134): `last` for levels 1-2; previous / one-before symbols
    // and the running drop sum for levels 3-4
    uint32_t last = 0, p1 = 0, p2 = 0, dl = 5;
    uint32_t ctx = 0, h = 0, used = 0;
    bool full = false;
    SfqQdModelT<LPC> m;
    m.zero();
    uint4 nv[G::NV];                                                      // the next model, as loaded (CH: header, two freq vectors, symbols)
    uint32_t nkey = 0;                                                    // CH: its key word
#pragma unroll
    for (int q = 0; q < G::NV; q++) nv[q] = make_uint4(0, 0, 0, 0);
    uint32_t ex1 = 0, ex2 = 0, ex3 = 0, key = 0;                          // CH: the prefixes of lanes 1..3 and the entry's key, in every lane
    bool claimed = false;                                                 // CH: the entry was taken by this visit (its key goes out with the header)
    // CH: fetch / unpack of an entry
#define SFQ_QD_FETCH(e_)                                                                                   \
    {                                                                                                      \
        const uint32_t *e__ = (e_);                                                                        \
        if (CH) {                                                                                          \
            /* (scalar loads: a 128-bit load pins its four words to an aligned register quad, and the register       \
               allocator then copies the first word out of the quad right behind the load - a first touch that waits   \
               out the whole DRAM latency before the model update meant to cover it) */                                \
            nv[0].x = __ldcg(e__); nv[0].y = __ldcg(e__ + 1); nv[0].z = __ldcg(e__ + 2); nv[0].w = __ldcg(e__ + 3);    \
            nkey = __ldcg(e__ + 4);                                                                        \
            nv[1] = __ldcg(reinterpret_cast<const uint4 *>(e__ + 8 + 8 * l8));                             \
            nv[2] = __ldcg(reinterpret_cast<const uint4 *>(e__ + 12 + 8 * l8));                            \
            nv[3] = __ldcg(reinterpret_cast<const uint4 *>(e__ + 40 + 4 * l8));                            \
        } else sfq_qd_fetch<LPC>(nv, e__);                                                                 \
    }
#define SFQ_QD_UNPACK()                                                                                    \
    {                                                                                                      \
        if (CH) {                                                                                          \
            const uint32_t fw_[8] = {nv[1].x, nv[1].y, nv[1].z, nv[1].w, nv[2].x, nv[2].y, nv[2].z, nv[2].w}; \
            _Pragma("unroll")                                                                              \
            for (int k_ = 0; k_ < 8; k_++) m.pw[k_ % G::FW] = fw_[k_];                                    \
            m.sy[0] = nv[3].x; m.sy[G::SW > 1 ? 1 : 0] = nv[3].y; m.sy[G::SW > 2 ? 2 : 0] = nv[3].z; m.sy[G::SW > 3 ? 3 : 0] = nv[3].w; \
            m.hdr = nv[0].x; ex1 = nv[0].y; ex2 = nv[0].z; ex3 = nv[0].w; key = nkey;                      \
            m.excl = l8 == 1u ? ex1 : l8 == 2u ? ex2 : l8 == 3u ? ex3 : 0u;                                \
            claimed = false;                                                                               \
        } else m.unpack(nv);                                                                               \
    }
    if (live) SFQ_QD_FETCH(tab)
    bool fresh = live;                                                    // nv holds a model not yet unpacked
    bool chk = live && !dense;

    while (__any_sync(FULL, live)) {
        if (CH) {
            // `fresh` as the compiler cannot see through it: the unpack below is partly plain copies of the loaded words, and
            // with the condition known to be the one the fetch ran under, those copies are resolved right behind the loads
            // at the bottom of the previous step - the first touch of the loaded registers, which then waits out the whole
            // DRAM latency before the model update that was meant to cover it (measured: long-scoreboard stalls x2.2)
            const bool fr = __shfl_sync(FULL, (uint32_t)fresh, lane) != 0u;
            if (fr) SFQ_QD_UNPACK()
        } else if (fresh) SFQ_QD_UNPACK()
        // ---------------------------------------------------------------- the entry of `ctx` (hash probe)
        // An entry is in use once its total is non-zero (every update adds to it, and every lane holds the
        // total); lane 0 keeps the 16-bit key in its otherwise unused prefix word.
        if (__any_sync(FULL, chk)) {
            uint32_t probes = 0;
            for (;;) {
                const bool occ = (m.hdr & 0x3fffffu) != 0u;
                const unsigned vb = __ballot_sync(FULL, chk && occ && (CH ? key != ctx : (l8 == 0 && m.excl != ctx)));
                const bool bad = CH ? (chk && occ && key != ctx) : (bool)((vb >> osh) & 1u);
                if (chk && !occ) {                                        // first visit of this context in the chunk
                    if (used + 1u >= nent) { full = true; live = false; }
                    else { used++; if (CH) { key = ctx; claimed = true; } else if (l8 == 0) m.excl = ctx; }
                    chk = false;
                } else if (chk && bad) {                                  // somebody else's entry: walk on
                    h = h + 1u == nent ? 0u : h + 1u;
                    if (++probes > nent) { full = true; live = false; chk = false; }
                    else { SFQ_QD_FETCH(tab + (size_t)h * 64u) SFQ_QD_UNPACK() }
                } else chk = false;
                if (!vb) break;
            }
        }
        m.derive();
        const bool act = live;
        if (SPEC) {
            // The next context is only known once this symbol is: its model is requested below, after the search, and its
            // DRAM latency sits on the chain.  A quality most often repeats its predecessor, and for that case the next
            // context is known NOW (qlts.hpp:52-74 with b = p1: same symbol twice, no drop): ask L2 for that entry's
            // line before the search starts.  A wrong guess costs a line of bandwidth and nothing else.
            const uint32_t guess = level >= 3u ? ((p1 | (p1 << 6) | (1u << 12) | ((dl >> 3 < 7u ? dl >> 3 : 7u) << 13)) & 0xffffu)
                                               : ((p1 | (last << 6)) & lmask);
            const uint32_t hg = dense ? guess : __umulhi(guess * 2654435761u, nent);
            if (act) sfq_prefetch(tab + (size_t)hg * 64u);
        }

        // ---------------------------------------------------------------- Log64Ranger::get (log64_ranger.hpp:114-138)
        const uint32_t tot = m.hdr & 0x3fffffu, count = m.hdr >> 24;
        const uint32_t rr = rc.range / (tot + 64u);
        const uint32_t e0 = (l8 ? m.excl : 0u) + SPL * l8;
        uint32_t cs[SPL];                                                  // cumulative (freq + 1) up to and including slot k
        cs[0] = e0 + m.f[0] + 1u;
#pragma unroll
        for (int k = 1; k < (int)SPL; k++) cs[k] = cs[k - 1] + m.f[k] + 1u;
        // (a corrupt stream can leave code >= 2^32: every comparison below is then true, as with the reference's 64-bit quotient)
        const uint32_t code32 = (uint32_t)(rc.code >> 32) ? 0xffffffffu : (uint32_t)rc.code;
        const unsigned sel = (__ballot_sync(FULL, code32 < cs[SPL - 1] * rr) >> osh) & ((1u << LPC) - 1u);
        const uint32_t hl = sel ? (uint32_t)(__ffs(sel) - 1) : (uint32_t)(LPC - 1);      // no lane: corrupt stream, last slot
        uint32_t cb, hk, pack;
        {   // binary search of code among this lane's thresholds cs[0..SPL-2] (meaningful in lane hl)
            uint32_t idx = 0, lo = e0, hi = cs[SPL - 1];
            // one level of the search: the candidate is cs[idx + STEP - 1] (all indices compile-time constants)
#define SFQ_QD_LEVEL(STEP)                                                                                   \
            {                                                                                                \
                uint32_t cand = cs[(STEP) - 1];                                                              \
                _Pragma("unroll")                                                                            \
                for (int b_ = 2 * (STEP); b_ < (int)SPL; b_ += 2 * (STEP)) cand = idx == (uint32_t)b_ ? cs[b_ + (STEP) - 1] : cand; \
                const bool g_ = code32 >= cand * rr;                                                         \
                lo = g_ ? cand : lo; hi = g_ ? hi : cand; idx += g_ ? (uint32_t)(STEP) : 0u;                  \
            }
            if (SPL == 16) SFQ_QD_LEVEL(8)
            SFQ_QD_LEVEL(4)
            SFQ_QD_LEVEL(2)
            SFQ_QD_LEVEL(1)
#undef SFQ_QD_LEVEL
            hk = idx;
            cb = lo;
            pack = (hi - lo - 1u) | ((m.sym_byte(hk) ^ (SPL * l8 + hk)) << 16) | (hk << 24);
        }
        cb = __shfl_sync(FULL, cb, hl, LPC);                                 // the chosen slot, from its lane
        pack = __shfl_sync(FULL, pack, hl, LPC);
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
#define SFQ_QD_BC32(x) { const uint32_t v_ = __shfl_sync(FULL, (uint32_t)(x), 0, LPC); if (esc) x = v_; }
#define SFQ_QD_BC64(x) { const uint32_t lo_ = __shfl_sync(FULL, (uint32_t)(x), 0, LPC), hi_ = __shfl_sync(FULL, (uint32_t)((uint64_t)(x) >> 32), 0, LPC); \
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
            SFQ_QD_FETCH(tab + (size_t)hn * 64u)
        }

        // ---------------------------------------------------------------- update_freq (log64_ranger.hpp:69-87)
        const uint32_t slot = SPL * hl + hk;
        bool skip = false, wide = false, halved = false, repack = false;
        uint32_t fn = f, tot2 = tot;
        if (__any_sync(FULL, act && f > 65472u - 6u)) {
            const bool hv = act && f > 65472u - 6u;
            if (hv && slot == 0u && f + 20u > tot) skip = true;               // saturated front slot: no update at all
            const bool dohalve = hv && !skip;
            uint32_t ls = 0;
#pragma unroll
            for (int k = 0; k < (int)SPL; k++) { if (dohalve) m.f[k] >>= 1; ls += m.f[k]; }
            uint32_t inc = ls, tmp;
#pragma unroll
            for (int d = 1; d < LPC; d <<= 1) { tmp = __shfl_up_sync(FULL, inc, d, LPC); if (l8 >= (uint32_t)d) inc += tmp; }
            const uint32_t total = __shfl_sync(FULL, inc, LPC - 1, LPC);
            if (CH) {                                                        // every lane keeps all three prefixes
                const uint32_t p1_ = __shfl_sync(FULL, inc - ls, 1, LPC), p2_ = __shfl_sync(FULL, inc - ls, 2, LPC), p3_ = __shfl_sync(FULL, inc - ls, 3, LPC);
                if (dohalve) { ex1 = p1_; ex2 = p2_; ex3 = p3_; }
            }
            if (dohalve) {
                if (l8) m.excl = inc - ls;
                tot2 = total;
                fn = f >> 1;
                wide = true; halved = true; repack = true;
            }
        }
        const bool upd = act && !skip;
        const bool mine = upd && l8 == hl;
        if (upd) {
            fn += 6u;
            tot2 += 6u;
            if (l8 > hl) m.excl += 6u;
            if (CH) { ex1 += hl < 1u ? 6u : 0u; ex2 += hl < 2u ? 6u : 0u; ex3 += hl < 3u ? 6u : 0u; }
        }
        if (mine) { if (halved) m.set_freq(hk, fn); else m.add_packed(hk, 6u); }      // (f[] of this lane is stale for slot hk from here on, unless halved)
        uint32_t cnt2 = count;
        const bool cand = upd && slot != 0u;                                 // `++count` is not evaluated for slot 0
        if (cand) cnt2 = (count + 1u) & 0xffu;
        const bool swapc = cand && (cnt2 & 0xfu) == 0u;
        bool syms_dirty = false;
        if (__any_sync(FULL, swapc)) {                                       // maybe swap slot with slot-1 (down_level)
            const uint32_t lfl = __shfl_up_sync(FULL, m.f[SPL - 1], 1, LPC);   // left neighbour's last slot
            const uint32_t lbl = __shfl_up_sync(FULL, m.sy[G::SW - 1] >> 24, 1, LPC);
            // lane hl works out the neighbour slot and the verdict, the group hears it
            uint32_t fprev, bprev;
            if (hk == 0u) { fprev = lfl; bprev = lbl; }
            else { fprev = m.freq_at(hk - 1u); bprev = m.sym_byte(hk - 1u); }
            uint32_t verdict = (swapc && fn > fprev ? 1u : 0u) | (fprev << 1) | (bprev << 24);
            verdict = __shfl_sync(FULL, verdict, hl, LPC);
            if (verdict & 1u) {
                fprev = (verdict >> 1) & 0xffffu; bprev = verdict >> 24;
                const uint32_t symprev = bprev ^ (slot - 1u);
                const uint32_t byte_lo = (ssym ^ (slot - 1u)) & 0xffu;        // slot-1 now holds this symbol ...
                const uint32_t byte_hi = (symprev ^ slot) & 0xffu;            // ... and slot the neighbour's
                if (CH && hk == 0u) {                                         // the slot moved into the lane before: that lane's prefix is unchanged, lane hl's grows
                    const uint32_t d_ = fn - fprev;
                    ex1 += hl == 1u ? d_ : 0u; ex2 += hl == 2u ? d_ : 0u; ex3 += hl == 3u ? d_ : 0u;
                }
                if (l8 == hl) {
                    if (hk == 0u) {
                        m.f[0] = fprev;
                        m.set_sym_byte(0u, byte_hi);
                        m.excl += fn - fprev;                                 // (hl >= 1 here: lane 0's word keeps the key)
                    } else {
                        m.set_freq(hk - 1u, fn); m.set_freq(hk, fprev);
                        m.set_sym_byte(hk - 1u, byte_lo); m.set_sym_byte(hk, byte_hi);
                    }
                    syms_dirty = true; repack = true;
                }
                if (hk == 0u && l8 + 1u == hl) {
                    m.f[SPL - 1] = fn;
                    m.set_sym_byte(SPL - 1u, byte_lo);
                    syms_dirty = true; wide = true; repack = true;
                }
            }
        }
        if (repack) m.pack();                // rare: a halving or a swap rewrote f[] (every slot of it is current in those lanes)
        if (act) {
            m.hdr = tot2 | (cnt2 << 24);
            uint32_t *e = tab + (size_t)h * 64u;
            if (CH) {
                if (wide || mine) {
                    uint32_t *fe = e + 8 + 8 * l8;
                    *reinterpret_cast<uint4 *>(fe) = make_uint4(m.pw[0], m.pw[1 % G::FW], m.pw[2 % G::FW], m.pw[3 % G::FW]);
                    *reinterpret_cast<uint4 *>(fe + 4) = make_uint4(m.pw[4 % G::FW], m.pw[5 % G::FW], m.pw[6 % G::FW], m.pw[7 % G::FW]);
                }
                if (syms_dirty) *reinterpret_cast<uint4 *>(e + 40 + 4 * l8) = make_uint4(m.sy[0], m.sy[G::SW > 1 ? 1 : 0], m.sy[G::SW > 2 ? 2 : 0], m.sy[G::SW > 3 ? 3 : 0]);
                if (l8 == 0) {
                    *reinterpret_cast<uint4 *>(e) = make_uint4(m.hdr, ex1, ex2, ex3);
                    if (claimed) e[4] = key;
                }
                claimed = false;
            } else {
                if (wide || mine) m.store_freqs(e);
                if (syms_dirty) m.store_syms(e);
                m.store_hdr(e);
            }
        }
        if (CH) __syncwarp();            // the header sector is written by lane 0 and read by all four lanes: order the warp's stores before its later loads
        if (fresh) { h = hn; ctx = nctx; chk = !dense; }
    }
#undef SFQ_QD_OPEN
#undef SFQ_QD_FETCH
#undef SFQ_QD_UNPACK
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
