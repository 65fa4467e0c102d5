// Per-chunk coders: one routine per (chunk, main stream).  Each routine is the whole serial
// dependency chain of one adaptive range-coded stream; the kernels in sfq_kernels.cu run one
// such routine per thread (thousands of chunk-streams resident at once).
//
// A chunk is coded exactly as the reference would code it as a standalone file:
//   gen  + gen.Ns + gen.Nn              GenSave::save / GenLoad::load      gens.cpp:91-159, 200-249
//   qlt                                 QltSave::save_1/2/3 / load_1/2/3   qlts.cpp:74-136, 163-234
//   rec  + rec.x + usr.x/.x.q/.pfg/.pfq RecSave::save / RecLoad::load      recs.cpp:277-461
//                                       UsrSave::get_record/update         usrs.cpp:124-156, 303-390
#pragma once
#include "sfq_coder.cuh"

// Decoder-side per-record length tables: an oversized record (decoded verbatim by sfq_usr_decode_chunk) carries this bit in
// its three length entries; the model decoders see length 0 for it.
#define SFQ_BIG_BIT 0x80000000u
SFQ_HD uint32_t sfq_coded_len(uint32_t table_entry) { return (table_entry & SFQ_BIG_BIT) ? 0u : table_entry; }

// Geometry of record r of a chunk, read from the line-start table built by the scan kernel.
// line_start[L] = offset of the first byte of line L; line_start[nlines] = end of text.
struct SfqRecView {
    const uint8_t *hdr;  uint32_t hlen;     // header without '@' and '\n'
    const uint8_t *seq;  uint32_t llen;     // coded bases (SOLiD prefix stripped); 0 for an oversized record
    const uint8_t *qual; uint32_t qlen;     // coded qualities (SOLiD prefix stripped); 0 for an oversized record
    uint32_t plus_len;                      // '+' line length including the '+'
    uint8_t pf_gen, pf_qlt;                 // SOLiD prefix chars (usrs.cpp:324-329,358-363)
    bool big;                               // oversized (usrs.hpp:34-36): not coded by the models; raw_llen / raw_qlen = what the lines measure
    uint32_t raw_llen, raw_qlen;
};
SFQ_HD SfqRecView sfq_rec_view(const uint8_t *text, const uint64_t *ls, uint64_t line0, uint32_t r, uint32_t solid) {
    const uint64_t *l = ls + line0 + 4ull * r;
    SfqRecView v;
    v.hdr = text + l[0] + 1;  v.hlen = (uint32_t)(l[1] - l[0] - 2);
    uint32_t sl = (uint32_t)(l[2] - l[1] - 1), ql = (uint32_t)(l[4] - l[3] - 1);
    v.plus_len = (uint32_t)(l[3] - l[2] - 1);
    v.seq = text + l[1];  v.qual = text + l[3];
    v.pf_gen = 0; v.pf_qlt = 0;
    if (solid) {
        // usrs.cpp:324-329: the prefix byte is consumed unconditionally, then the line is scanned.
        v.pf_gen = v.seq[0]; v.pf_qlt = v.qual[0];
        v.seq++; v.qual++;
        sl = sl ? sl - 1 : 0; ql = ql ? ql - 1 : 0;
    }
    v.raw_llen = sl; v.raw_qlen = ql;
    v.big = v.hlen >= SFQ_MAX_ID_LLEN - 1 || sl >= SFQ_MAX_GN_LLEN - 1 || ql >= SFQ_MAX_GN_LLEN - 1;
    v.llen = v.big ? 0u : sl; v.qlen = v.big ? 0u : ql;
    return v;
}

// gencodes[] (gens.cpp:72-77) without a table or a jump: ACGT/acgt/0123 -> 0..3, N n . -> 4, else 0x10.
SFQ_HD uint32_t sfq_gencode(uint8_t c) {
    const uint32_t u = c & 0xDFu;                         // fold case (also maps '0'..'3' -> 0x10..0x13)
    // A=0x41 C=0x43 G=0x47 T=0x54: bits 1..2 give 0,1,3,2 -> swap the last two
    uint32_t n = (u >> 1) & 3u;
    n ^= n >> 1;
    const bool acgt = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
    const bool digit = (uint32_t)(c - '0') < 4u;
    const bool isn = (u == 'N') | (c == '.');
    return acgt ? n : digit ? (uint32_t)(c - '0') : isn ? 4u : 0x10u;
}
SFQ_HD uint32_t sfq_gen_mask(int level) {           // gens.hpp:43-53,74-82
    return level <= 1 ? (1u << 18) - 1 : level == 2 ? (1u << 22) - 1 : level == 3 ? (1u << 24) - 1 : (1u << 26) - 1;
}

// ============================================================================ gen: encode
// Run-ahead cursor: an encoder's contexts depend on the input alone, so a second cursor walks the
// bases SFQ_GEN_AHEAD symbols ahead of the coder and prefetches the table slot each one will need;
// the coder then finds its slot in cache instead of paying an HBM round trip per base.
#define SFQ_GEN_AHEAD 48u
struct SfqGenCursor {
    const uint8_t *text; const uint64_t *ls; uint64_t line0;
    uint32_t nrec, solid, mask, r, i, llen, last;
    SfqReader rd;
    SFQ_HD void open_record() {
        while (r < nrec) {
            const SfqRecView v = sfq_rec_view(text, ls, line0, r, solid);
            llen = v.llen; i = 0; last = 0x007616c7u;
            if (llen) { rd.seek(v.seq); return; }
            r++;
        }
        llen = 0;
    }
    SFQ_HD void init(const uint8_t *t, const uint64_t *l, const SfqChunkMeta *m, uint32_t msk) {
        text = t; ls = l; line0 = m->line0; nrec = m->nrec; solid = m->solid; mask = msk; r = 0;
        open_record();
    }
    // advance up to `count` bases, prefetching each one's slot
    SFQ_HD void run(const SfqGenTable &tab, uint32_t count) {
        while (count && r < nrec) {
            uint32_t n = sfq_gencode(rd.next());
            if (n > 3) n = 0;
            last &= mask;
            tab.prefetch(last);
            last = (last << 2) | n;
            count--;
            if (++i == llen) { r++; open_record(); }
        }
    }
};

SFQ_HDN void sfq_gen_encode_chunk(const uint8_t *text, const uint64_t *ls, SfqChunkMeta *meta, int level,
                                  void *table_mem, uint32_t hbits, uint32_t *pwpool,
                                  uint8_t *arena, SfqArena *ar) {
    SfqEnc rc;
    rc.start(arena + ar->off[SFQ_S_GEN], ar->cap[SFQ_S_GEN]);
    SfqXSave xns, xnn;
    xns.init(pwpool, SFQ_X_NS, arena + ar->off[SFQ_S_GEN_NS], ar->cap[SFQ_S_GEN_NS]);
    xnn.init(pwpool, SFQ_X_NN, arena + ar->off[SFQ_S_GEN_NN], ar->cap[SFQ_S_GEN_NN]);
    SfqGenTable tab;
    tab.init(table_mem, hbits, level <= 1);
    const uint32_t mask = sfq_gen_mask(level);
    const uint32_t solid = meta->solid;
    uint64_t genofs = 0, ns_index = 0, nn_index = 0;     // g_genofs_count, m_last.*  (gens.hpp:56-60)
    uint8_t n_byte = 0;
    uint32_t status = SFQ_OK, status_arg = 0;
    SfqGenCursor ahead;
    ahead.init(text, ls, meta, mask);
    ahead.run(tab, SFQ_GEN_AHEAD);

    for (uint32_t r = 0; r < meta->nrec && status == SFQ_OK; r++) {
        const SfqRecView v = sfq_rec_view(text, ls, meta->line0, r, solid);
        uint32_t last = 0x007616c7u;                                           // gens.cpp:139
        SfqReader rs, rq;
        if (v.llen) rs.seek(v.seq);
        if (v.qlen) rq.seek(v.qual);
        for (uint32_t i = 0; i < v.llen; i++) {
            if ((i & 15u) == 0) ahead.run(tab, 16);
            const uint8_t g = rs.next();
            const uint8_t q = i < v.qlen ? rq.next() : (uint8_t)40;            // gens.cpp:153
            uint32_t n = sfq_gencode(g);
            const bool bad_q = (q == '!');
            bool bad_n = false;
            if (n > 3) {
                if (n > 4) { status = SFQ_E_BASE; status_arg = g; break; }
                bad_n = true; n = 0;
            }
            genofs++;
            if (bad_n || bad_q) {                                              // gens.cpp:91-114
                if (!bad_n) { xnn.put(genofs - nn_index); nn_index = genofs; }
                else {
                    if (!n_byte) n_byte = g;
                    if (g != n_byte) { status = SFQ_E_NBYTE; status_arg = g; break; }
                    if (!bad_q) { xns.put(genofs - ns_index); ns_index = genofs; }
                }
            }
            last &= mask;
            uint32_t fv;
            const uint32_t slot = tab.find(last, fv);
            if (slot == 0xFFFFFFFFu) { status = SFQ_E_TABLE; break; }
            tab.store(slot, last, sfq_b2_put(fv, rc, n));
            last = (last << 2) | n;
        }
    }
    rc.finish();
    bool ovf = rc.out.overflow();
    ar->size[SFQ_S_GEN] = rc.out.n;
    ar->size[SFQ_S_GEN_NS] = xns.close(ovf);
    ar->size[SFQ_S_GEN_NN] = xnn.close(ovf);
    if (status == SFQ_OK && ovf) status = SFQ_E_CAP;
    meta->n_byte = (n_byte && n_byte != 'N') ? n_byte : 0;                     // gens.cpp:103-104
    meta->g_used = tab.used;
    if (status != SFQ_OK && meta->status == SFQ_OK) { meta->status = status; meta->status_arg = status_arg; }
}

// ============================================================================ gen: decode
// Writes the base line of every record into `bases` (record r at boff[r], llen[r] bytes).  The
// "quality '!' means N" rule (GenLoad::normalize_gen, gens.cpp:200-213) needs the record's decoded
// qualities, which another thread is producing concurrently, so it is applied afterwards by the
// assemble kernel: positions taken from gen.Nn (a real base under a '!' quality) are flagged with
// bit 7 here so the rule skips them; gen.Ns positions get the N byte right away.
//
// This loop is one serial chain of ~440 k steps per chunk, so it is written for few instructions and
// one memory round trip per base:
//   * contexts live in buckets of four 64-bit slots (one 32-byte sector): a lookup is one sector load
//     (two 16-byte loads), key compare and first-empty search in registers;
//   * range / totFreq (coder.hpp:84) multiplies by a reciprocal from a 1021-entry table (totFreq <= 1020);
//   * the symbol search compares `code` with cumFreq * range instead of dividing code by range
//     (coder.hpp:85, base2_ranger.hpp:91-98): floor(code / r) < c  <=>  code < c * r;
//   * the two exception lists are watched through one 32-bit "next exception" position.
SFQ_HD uint32_t sfq_umulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// floor(2^32 / d) for d >= 2
SFQ_HD uint32_t sfq_recip_u32(uint32_t d) {
    const uint32_t q = 0xffffffffu / d;
    return q + ((0xffffffffu - q * d) == d - 1u ? 1u : 0u);
}
// floor(n / d) from inv = floor(2^32 / d)
SFQ_HD uint32_t sfq_div_recip(uint32_t n, uint32_t d, uint32_t inv) {
    const uint32_t q = sfq_umulhi(n, inv);
    return q + ((n - q * d) >= d ? 1u : 0u);
}
#define SFQ_B2_LUT 1024u               // reciprocals of totFreq 0..1023 (entries 0,1 unused)
SFQ_HD void sfq_b2_lut_fill(uint32_t *lut, uint32_t first, uint32_t step) {
    for (uint32_t t = first; t < SFQ_B2_LUT; t += step) lut[t] = t < 2 ? 0xffffffffu : sfq_recip_u32(t);
}

// Decoder-side base-context table: `nb` buckets of 4 slots (nb a multiple of 4), slot = (ctx+1) << 32 | freq[4];
// four consecutive buckets share one 128-byte line.
// The LINE of a context is chosen by its grandparent (ctx >> 4, i.e. without the two newest bases) and the
// bucket inside the line by the older of those two bases: the line of base i+2 and the bucket of base i+1
// therefore depend only on the context of base i.  While base i is being decoded the line of base i+2 is
// prefetched into L2 and the bucket of base i+1 is copied to shared memory, so the table's HBM latency is
// spread over two links of the chain and the bucket copy itself is an L2 hit.
// (Level 1 keeps the direct 2^18 x u32 table, stored ^0x03030303.)
struct SfqGenBuckets {
    uint64_t *slots;
    uint32_t nb, nl, used, dense, ahead2;
    SFQ_HD void init(void *mem, uint32_t nbuckets, bool is_dense) { slots = (uint64_t *)mem; nb = nbuckets & ~3u; nl = nb >> 2; used = 0; dense = is_dense; ahead2 = 1; }
    SFQ_HD uint32_t home(uint32_t ctx) const { return 4u * sfq_umulhi((ctx >> 4) * 2654435761u, nl) + ((ctx >> 2) & 3u); }
    // home bucket of whichever context follows `ctx` (mask = context mask of the level)
    SFQ_HD uint32_t next_home(uint32_t ctx, uint32_t mask) const { return 4u * sfq_umulhi(((ctx >> 2) & (mask >> 4)) * 2654435761u, nl) + (ctx & 3u); }
    // line of whichever context comes two bases after `ctx`
    SFQ_HD uint32_t line_after2(uint32_t ctx, uint32_t mask) const { return sfq_umulhi((ctx & (mask >> 4)) * 2654435761u, nl); }
    SFQ_HD void prefetch_line(uint32_t line) const {
        if (!ahead2) return;
        const uint64_t *p = slots + 16ull * line;
        sfq_prefetch(p); sfq_prefetch(p + 4); sfq_prefetch(p + 8); sfq_prefetch(p + 12);
    }
};
SFQ_HD void sfq_ld_bucket(const uint64_t *p, uint32_t (&k)[4], uint32_t (&v)[4]) {
#if defined(__CUDA_ARCH__)
    const uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + 2);
    v[0] = a.x; k[0] = a.y; v[1] = a.z; k[1] = a.w; v[2] = b.x; k[2] = b.y; v[3] = b.z; k[3] = b.w;
#else
    for (int j = 0; j < 4; j++) { v[j] = (uint32_t)p[j]; k[j] = (uint32_t)(p[j] >> 32); }
#endif
}
// Looks for `key` among the four slots held in registers: true if it is there or a slot is free
// (j = slot index, fv = its frequencies or the start state); slots fill front to back.
SFQ_HD bool sfq_bucket_pick(const uint32_t (&k)[4], const uint32_t (&v)[4], uint32_t key, uint32_t &j, uint32_t &fv, bool &hit) {
    const bool m0 = k[0] == key, m1 = k[1] == key, m2 = k[2] == key, m3 = k[3] == key;
    hit = m0 | m1 | m2 | m3;
    const uint32_t nfree = (k[0] == 0) + (k[1] == 0) + (k[2] == 0) + (k[3] == 0);
    j = hit ? (m0 ? 0u : m1 ? 1u : m2 ? 2u : 3u) : 4u - nfree;
    fv = m0 ? v[0] : m1 ? v[1] : m2 ? v[2] : m3 ? v[3] : 0x03030303u;
    return hit | (nfree != 0);
}
// Look-ahead of one bucket.  On the device the 32 bytes travel global -> shared with cp.async: that
// copy is not tracked by the register scoreboard, so it stays in flight across the divergent branches of
// the coder's renormalisation loop (a plain early load is waited for at the first such branch) and is
// only collected when the next base needs it.  `slot` = this thread's two 16-byte cells in shared memory.
struct SfqStage {
    void *cell0 = nullptr, *cell1 = nullptr;
    SFQ_HD void request(const uint64_t *bucket) const {
#if defined(__CUDA_ARCH__)
        const unsigned a0 = (unsigned)__cvta_generic_to_shared(cell0), a1 = (unsigned)__cvta_generic_to_shared(cell1);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a0), "l"(bucket));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a1), "l"(bucket + 2));
        asm volatile("cp.async.commit_group;");
#else
        (void)bucket;
#endif
    }
    SFQ_HD void collect(const uint64_t *bucket, uint32_t (&k)[4], uint32_t (&v)[4]) const {
#if defined(__CUDA_ARCH__)
        (void)bucket;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const uint4 a = *reinterpret_cast<const uint4 *>(cell0), b = *reinterpret_cast<const uint4 *>(cell1);
        v[0] = a.x; k[0] = a.y; v[1] = a.z; k[1] = a.w; v[2] = b.x; k[2] = b.y; v[3] = b.z; k[3] = b.w;
#else
        sfq_ld_bucket(bucket, k, v);
#endif
    }
};

// Byte source of a decoder: aligned 8-byte loads, one word ahead of the coder so that a refill never
// waits on memory; bytes past the end read as 0 (filer.hpp:94-97).
struct SfqByteSrc {
    const uint8_t *p;      // address of the word after `ahead`
    const uint8_t *end;
    uint64_t word, ahead;
    uint32_t left;         // unread bytes in `word`
    SFQ_HD uint64_t fetch() {
        uint64_t w = 0;
        if (p + 8 <= end) w = *reinterpret_cast<const uint64_t *>(p);
        else for (int k = 0; k < 8; k++) if (p + k < end) w |= (uint64_t)p[k] << (8 * k);
        p += 8;
        return w;
    }
    // the stream starts at any byte offset; loads are aligned, so the first word is entered part-way
    SFQ_HD void start(const uint8_t *buf, uint32_t size) {
        end = buf + size;
        const uint32_t mis = (uint32_t)((uintptr_t)buf & 7u);
        p = buf - mis;
        word = 0;
        for (uint32_t k = mis; k < 8; k++) if (p + k < end) word |= (uint64_t)p[k] << (8 * k);
        word >>= 8 * mis;
        p += 8;
        left = 8 - mis;
        ahead = fetch();
    }
    SFQ_HD uint32_t next() {
        if (left == 0) {
            word = ahead; ahead = fetch(); left = 8;
            if (((uintptr_t)p & 127u) == 0 && p + 256 <= end) sfq_prefetch_l1(p + 128);     // the stream is read once, front to back
        }
        const uint32_t c = (uint32_t)word & 0xffu;
        word >>= 8;
        left--;
        return c;
    }
};

// The two exception lists of the base stream (gen.Ns: an N whose quality is not '!'; gen.Nn: a real base
// under a '!' quality; gens.cpp:91-114, 188-189, 244-247) are sparse, so they are applied to the chunk's
// decoded base plane afterwards instead of being polled once per base: gen.Nn positions get bit 7 set
// (the assemble kernel's "quality '!' means N" rule skips them), gen.Ns positions get the N byte.
// Where both lists name a position the reference's first test (gen.Nn) wins, hence gen.Nn goes first.
SFQ_HDN void sfq_gen_apply_exceptions(const uint8_t *in, const uint32_t *ssize, const uint64_t *soff,
                                      const SfqChunkMeta *meta, uint32_t *pwpool, uint8_t *plane) {
    const uint64_t nb = meta->nbases;
    const uint8_t n_byte = meta->n_byte ? meta->n_byte : (uint8_t)'N';          // gens.cpp:169
    for (int pass = 0; pass < 2; pass++) {
        SfqXLoad x;
        if (pass == 0) x.init(pwpool, SFQ_X_NN, in + soff[SFQ_S_GEN_NN], ssize[SFQ_S_GEN_NN]);
        else x.init(pwpool, SFQ_X_NS, in + soff[SFQ_S_GEN_NS], ssize[SFQ_S_GEN_NS]);
        uint64_t index = 0;
        for (;;) {
            const uint64_t d = x.get();
            if (d == 0) break;                       // terminator, or an absent stream (xfile.cpp:90-93)
            const uint64_t nxt = index + d;
            if (nxt <= index || nxt > nb) break;     // never reached by the reference's running base count
            index = nxt;
            uint8_t *c = plane + (index - 1);
            if (pass == 0) *c |= 0x80u;
            else if (!(*c & 0x80u)) *c = n_byte;
        }
    }
}

template <bool DENSE>
SFQ_HD uint32_t sfq_gen_decode_loop(SfqByteSrc &src, SfqGenBuckets &tab, uint32_t mask, uint32_t alpha,
                                    const SfqChunkMeta *meta, const uint32_t *llen_tab,
                                    const uint64_t *boff_tab, uint8_t *bases, const uint32_t *lut, SfqStage stage) {
    uint32_t *dtab = reinterpret_cast<uint32_t *>(tab.slots);
    uint64_t low = 0, code = 0;
    uint32_t range = 0xFFFFFFFFu;
    for (int i = 0; i < 8; i++) code = (code << 8) | src.next();               // coder.hpp:44-48
    const uint32_t nrec = meta->nrec;
    for (uint32_t r = 0; r < nrec; r++) {
        const uint32_t llen = sfq_coded_len(llen_tab[r]);
        uint8_t *g = bases + boff_tab[r];
        uint32_t last = 0x007616c7u;
        uint32_t bk = DENSE ? 0u : tab.home(last & mask), bkn = 0;
        uint32_t k[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
        uint32_t pj = 4, pkey = 0, pfv = 0;     // slot of the current bucket written after it was loaded (4 = none)
        if (!DENSE && llen) {
            stage.request(tab.slots + 4ull * bk);
            if (llen > 1) tab.prefetch_line(tab.next_home(last & mask, mask) >> 2);
        }
        for (uint32_t i = 0; i < llen; i++) {
            const uint32_t ctx = last & mask;
            uint32_t fv;
            uint64_t *slot = nullptr;
            if (DENSE) fv = dtab[ctx] ^ 0x03030303u;
            else {
                // the bucket was requested before the previous base stored its slot: if that store went into
                // this very bucket, the registers are one update behind
                stage.collect(tab.slots + 4ull * bk, k, v);
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) if (q == pj) { k[q] = pkey; v[q] = pfv; }
                uint32_t j; bool hit;
                if (sfq_bucket_pick(k, v, ctx + 1u, j, fv, hit)) slot = tab.slots + 4ull * bk + j;
                else {                      // home bucket full of other contexts: walk on (rare)
                    uint32_t b2 = bk;
                    for (uint32_t probes = 0; !slot && probes < tab.nb; probes++) {
                        b2 = b2 + 1u == tab.nb ? 0u : b2 + 1u;
                        uint32_t k2[4], v2[4];
                        sfq_ld_bucket(tab.slots + 4ull * b2, k2, v2);
                        if (sfq_bucket_pick(k2, v2, ctx + 1u, j, fv, hit)) slot = tab.slots + 4ull * b2 + j;
                    }
                    if (!slot) return SFQ_E_TABLE;
                }
                tab.used += hit ? 0u : 1u;
                // the bucket of the next base does not depend on what this base decodes to: request it now,
                // use it in the next iteration
                bkn = tab.next_home(ctx, mask);
                if (i + 1 < llen) stage.request(tab.slots + 4ull * bkn);      // (never two copies in flight into the same cells)
                if (i + 2 < llen) tab.prefetch_line(tab.line_after2(ctx, mask));
            }
            const uint32_t f0 = fv & 0xff, f1 = (fv >> 8) & 0xff, f2 = (fv >> 16) & 0xff, f3 = fv >> 24;
            const uint32_t tot = f0 + f1 + f2 + f3;
            const uint32_t rr = sfq_div_recip(range, tot, lut[tot]);            // GetFreq: range /= tot
            uint32_t b, cr, f;
            if ((code >> 32) == 0) {
                const uint32_t c32 = (uint32_t)code;
                const uint32_t c1 = f0 * rr, c2 = c1 + f1 * rr, c3 = c2 + f2 * rr;
                const bool g1 = c32 >= c1, g2 = c32 >= c2, g3 = c32 >= c3;
                b = (uint32_t)g1 + (uint32_t)g2 + (uint32_t)g3;
                cr = g3 ? c3 : g2 ? c2 : g1 ? c1 : 0u;
                f = g3 ? f3 : g2 ? f2 : g1 ? f1 : f0;
            } else {                          // only a corrupt stream: behave like the reference's 64-bit divide
                const uint32_t prob = (uint32_t)(code / rr);
                if (prob < f0) { b = 0; cr = 0; f = f0; }
                else if (prob < f0 + f1) { b = 1; cr = f0 * rr; f = f1; }
                else if (prob < f0 + f1 + f2) { b = 2; cr = (f0 + f1) * rr; f = f2; }
                else { b = 3; cr = (f0 + f1 + f2) * rr; f = f3; }
            }
            low += cr;                         // Decode (coder.hpp:88-102)
            code -= cr;
            range = rr * f;
            while (range < SFQ_RC_TOP) {
                if ((low ^ (low + range)) & (0xffULL << 56))
                    range = (((uint32_t)low) | (SFQ_RC_TOP - 1)) - (uint32_t)low;
                code = (code << 8) | src.next();
                range <<= 8;
                low <<= 8;
            }
            fv = sfq_b2_update(fv, b);
            if (DENSE) dtab[ctx] = fv ^ 0x03030303u;
            else {
                *slot = ((uint64_t)(ctx + 1u) << 32) | fv;
                const uint32_t sidx = (uint32_t)(slot - tab.slots);
                pj = (sidx >> 2) == bkn ? (sidx & 3u) : 4u; pkey = ctx + 1u; pfv = fv;
                bk = bkn;
            }
            last = (last << 2) + b;
            g[i] = (uint8_t)(alpha >> (8 * b));  // exceptions: sfq_gen_apply_exceptions, afterwards
        }
    }
    return SFQ_OK;
}

SFQ_HDN void sfq_gen_decode_chunk(const uint8_t *in, const uint32_t *ssize, const uint64_t *soff,
                                  SfqChunkMeta *meta, int level, void *table_mem, uint32_t nbuckets,
                                  uint32_t *pwpool, const uint32_t *llen_tab, const uint64_t *boff_tab,
                                  uint8_t *bases, const uint32_t *lut, SfqStage stage, uint32_t ahead2 = 1) {
    SfqByteSrc src;
    src.start(in + soff[SFQ_S_GEN], ssize[SFQ_S_GEN]);
    (void)pwpool;
    SfqGenBuckets tab;
    tab.init(table_mem, nbuckets, level <= 1);
    tab.ahead2 = ahead2;
    const uint32_t mask = sfq_gen_mask(level);
    const uint32_t alpha = meta->solid ? 0x33323130u : 0x54474341u;             // "0123" / "ACGT", gens.cpp:173-178
    const uint32_t status = level <= 1
        ? sfq_gen_decode_loop<true>(src, tab, mask, alpha, meta, llen_tab, boff_tab, bases, lut, stage)
        : sfq_gen_decode_loop<false>(src, tab, mask, alpha, meta, llen_tab, boff_tab, bases, lut, stage);
    if (status != SFQ_OK && meta->status == SFQ_OK) meta->status = status;
}

// ============================================================================ qlt
struct SfqQCtx {                                      // qlts.cpp:109-112
    uint32_t last, delta, di;
    uint8_t q1, q2;
    SFQ_HD void reset() { last = 0; delta = 5; di = 0; q1 = 0; q2 = 0; }
};
SFQ_HD uint32_t sfq_q_delta_ctx(uint32_t &delta, uint8_t q, uint8_t q1, uint8_t q2) {   // qlts.hpp:62-74
    if (q1 > q) delta += (uint32_t)(q1 - q);
    const uint32_t d = delta >> 3;
    return ((uint32_t)q | ((uint32_t)(q1 < q2 ? q2 : q1) << 6) | ((uint32_t)(q1 == q2) << 12)
            | ((d < 7u ? d : 7u) << 13)) & 0xFFFFu;
}
SFQ_HD void sfq_q_next(SfqQCtx &c, int level, uint8_t b) {
    if (level <= 1) { c.last = ((uint32_t)b | (c.last << 6)) & 0xFFFu; return; }        // qlts.hpp:52-54
    if (level == 2) { c.last = ((uint32_t)b | (c.last << 6)) & 0xFFFFu; return; }       // qlts.hpp:55-57
    if (++c.di & 1u) { c.last = sfq_q_delta_ctx(c.delta, b, c.q1, c.q2); c.q2 = b; }    // qlts.cpp:127-134
    else             { c.last = sfq_q_delta_ctx(c.delta, b, c.q2, c.q1); c.q1 = b; }
}

#define SFQ_QLT_AHEAD 48u
// Thread-serial form of the quality coder over a direct table.  The GPU runs the lane-cooperative form
// (sfq_qlt_group.cuh); this one is what the CPU emulation checks against the oracle and it defines
// the behaviour the cooperative form must reproduce.
SFQ_HDN void sfq_qlt_encode_chunk(const uint8_t *text, const uint64_t *ls, SfqChunkMeta *meta, int level,
                                  uint32_t *qtable, uint32_t *pwpool, uint8_t *arena, SfqArena *ar) {
    SfqEnc rc;
    rc.start(arena + ar->off[SFQ_S_QLT], ar->cap[SFQ_S_QLT]);
    SfqPower ex; ex.m = pwpool + (size_t)SFQ_PW_QEX * SFQ_PW_WORDS;
    const uint32_t solid = meta->solid;
    uint32_t extra_hi = 0;
    for (uint32_t r = 0; r < meta->nrec; r++) {
        const SfqRecView v = sfq_rec_view(text, ls, meta->line0, r, solid);
        SfqQCtx c; c.reset();
        SfqReader rq;
        if (v.qlen) rq.seek(v.qual);
        for (uint32_t i = 0; i < v.qlen; i++) {
            const uint8_t b = (uint8_t)(rq.next() - '!');
            SfqLog64 m; m.m = qtable + (size_t)c.last * SFQ_L64_WORDS;
            if (b < 63) m.put(rc, b);
            else { m.put(rc, 63); ex.put(rc, b); extra_hi++; }                  // qlts.cpp:120-125
            sfq_q_next(c, level, b);
        }
    }
    rc.finish();
    ar->size[SFQ_S_QLT] = rc.out.n;
    meta->extra_hi = extra_hi;
    if (rc.out.overflow() && meta->status == SFQ_OK) meta->status = SFQ_E_CAP;
}

SFQ_HDN void sfq_qlt_decode_chunk(const uint8_t *in, const uint32_t *ssize, const uint64_t *soff,
                                  SfqChunkMeta *meta, int level, uint32_t *qtable, uint32_t *pwpool,
                                  const uint32_t *qlen_tab, const uint64_t *qoff_tab, uint8_t *quals) {
    SfqDec rc;
    rc.start(in + soff[SFQ_S_QLT], ssize[SFQ_S_QLT]);
    SfqPower ex; ex.m = pwpool + (size_t)SFQ_PW_QEX * SFQ_PW_WORDS;
    for (uint32_t r = 0; r < meta->nrec; r++) {
        const uint32_t qlen = sfq_coded_len(qlen_tab[r]);
        uint8_t *q = quals + qoff_tab[r];
        SfqQCtx c; c.reset();
        for (uint32_t i = 0; i < qlen; i++) {
            SfqLog64 m; m.m = qtable + (size_t)c.last * SFQ_L64_WORDS;
            uint32_t b = m.get(rc);
            if (b == 63) b = ex.get(rc);                                        // qlts.cpp:206-208
            q[i] = (uint8_t)('!' + b);
            sfq_q_next(c, level, (uint8_t)b);
        }
    }
}

// ============================================================================ rec (+ usr framing)
enum { SFQ_ST_DGT = 0, SFQ_ST_DLT, SFQ_ST_STR, SFQ_ST_HGT, SFQ_ST_HLT, SFQ_ST_HGT_Z, SFQ_ST_HLT_Z,
       SFQ_ST_HGTC, SFQ_ST_HLTC, SFQ_ST_HGTC_Z, SFQ_ST_HLTC_Z, SFQ_ST_DGT_Z, SFQ_ST_DLT_Z };   // recs.cpp:160-189

struct SfqSpaceMap {                                  // recs.hpp:69-74
    uint16_t off[66], wln[66];
    uint8_t str[66];
    uint32_t len;
};
SFQ_HD bool sfq_is_dig(uint8_t c) { return c >= '0' && c <= '9'; }
SFQ_HD bool sfq_is_word(uint8_t c) {                  // isdigit||isalpha in the C locale, recs.cpp:139
    return sfq_is_dig(c) || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
}
// map_space (recs.cpp:140-157); returns false on > 64 separators.
SFQ_HD bool sfq_map_space(SfqSpaceMap &m, const uint8_t *p) {
    m.len = 0;
    m.off[0] = 0;
    for (uint32_t i = 0;; i++) {
        const uint8_t c = p[i];
        if (!sfq_is_word(c)) {
            m.wln[m.len] = (uint16_t)(i - m.off[m.len]);
            m.str[m.len++] = c;
            m.off[m.len] = (uint16_t)(i + 1);
            if (c == 0 || c == '\n') return true;
            if (m.len > 64) return false;
        }
    }
}
// What kind of number, if any, a header field is (the verdict of recs.cpp:192-262): one pass classifies the characters
// after an optional single leading zero, then the value is accumulated in the base the classes call for.
//   classes   0 = decimal digit, 1 = a-f, 2 = A-F, 3 = anything else
//   decimal   every character a digit and the previous value of the field was not hexadecimal (pctype != 2); a value
//             that wraps 64 bits on the way (value * 10 + digit < value) makes the field a string
//   hex       at most 16 characters over all, one letter case only; tried when the previous value was hexadecimal or
//             when the first non-digit is a hex letter
// Two leading zeros cannot be reproduced by the decoder's "%lld" / "%llx": string.
SFQ_HD uint32_t sfq_hex_class(uint8_t c) {
    return (uint32_t)(c - '0') < 10u ? 0u : (uint32_t)(c - 'a') < 6u ? 1u : (uint32_t)(c - 'A') < 6u ? 2u : 3u;
}
SFQ_HD uint32_t sfq_hex_value(uint8_t c) { return (c & 0x40u) ? (c & 7u) + 9u : (uint32_t)(c - '0'); }
SFQ_HD uint32_t sfq_numberwang(const uint8_t *p, uint32_t len, uint64_t &num, uint8_t pctype) {
    const bool lead0 = p[0] == '0';
    if (lead0 && p[1] == '0') return SFQ_ST_STR;
    const uint32_t from = lead0 ? 1u : 0u;
    uint32_t seen = 0, run = 0;                        // classes met; length of the leading run of digits
    for (uint32_t k = from; k < len; k++) {
        const uint32_t cls = sfq_hex_class(p[k]);
        if (!seen || seen == 1u) run += cls == 0u;     // still inside the leading digit run
        seen |= 1u << cls;
    }
    num = 0;
    if (pctype != 2) {
        uint64_t v = 0;
        for (uint32_t k = from; k < from + run; k++) {
            const uint64_t t = v * 10u + (uint64_t)(p[k] - '0');
            if (t < v) return SFQ_ST_STR;
            v = t;
        }
        if (from + run >= len) { num = v; return lead0 ? SFQ_ST_DGT_Z : SFQ_ST_DGT; }
        if (sfq_hex_class(p[from + run]) == 3u) return SFQ_ST_STR;
    }
    if (len > 16 || (seen & 8u) || (seen & 6u) == 6u) return SFQ_ST_STR;
    uint64_t v = 0;
    for (uint32_t k = from; k < len; k++) v = (v << 4) + sfq_hex_value(p[k]);
    num = v;
    return (seen & 4u) ? (lead0 ? SFQ_ST_HGTC_Z : SFQ_ST_HGTC) : (lead0 ? SFQ_ST_HGT_Z : SFQ_ST_HGT);
}

// Working state of one chunk's header coder.  On the device it lives in SHARED memory: as local arrays it sat in L1, and
// L1 is what the coder kernels' neighbours on the SM take away (k_gen_replay and k_qlt_model carve 150-220 KB of shared
// memory out of the 228 KB), so every byte of tokenising went to L2.  `hbuf` stages the current and the previous id line
// (lines longer than the stage are read in place).
#define SFQ_REC_STAGE 240u
struct SfqRecScratch {
    SfqSpaceMap smap[2];
    uint64_t cnumb[2][65];
    uint8_t ctype[2][65];
    uint8_t hbuf[2][SFQ_REC_STAGE + 16];
};
// the id line `src` (hlen characters and the newline the tokeniser stops at) -> stage, or `src` itself if it does not fit
SFQ_HD const uint8_t *sfq_rec_stage(uint8_t *stage, const uint8_t *src, uint32_t hlen) {
    if (hlen + 1u > SFQ_REC_STAGE) return src;
    for (uint32_t k = 0; k <= hlen; k++) stage[k] = src[k];
    return stage;
}

struct SfqFieldRangers {                              // RecBase::ranger_t, recs.hpp:42-46
    SfqPower type, str;
    SfqPowerU num;
    SFQ_HD void bind(uint32_t *pool, uint32_t field) {
        uint32_t *b = pool + (size_t)(SFQ_PW_REC_BASE + field * SFQ_PW_PER_FIELD) * SFQ_PW_WORDS;
        type.m = b; str.m = b + SFQ_PW_WORDS; num.m = b + 2 * SFQ_PW_WORDS;
    }
};

SFQ_HDN void sfq_rec_encode_chunk(const uint8_t *text, const uint64_t *ls, SfqChunkMeta *meta,
                                  uint32_t *pwpool, uint8_t *arena, SfqArena *ar, SfqRecScratch *scr) {
    SfqEnc rc;
    rc.start(arena + ar->off[SFQ_S_REC], ar->cap[SFQ_S_REC]);
    SfqXSave x_rec, x_llen, x_qlen, x_sgen, x_sqlt, x_lrec, x_lgen, x_lqlt;
    x_lrec.init(pwpool, SFQ_X_LREC, arena + ar->off[SFQ_S_USR_LREC], ar->cap[SFQ_S_USR_LREC]);
    x_lgen.init(pwpool, SFQ_X_LGEN, arena + ar->off[SFQ_S_USR_LGEN], ar->cap[SFQ_S_USR_LGEN]);
    x_lqlt.init(pwpool, SFQ_X_LQLT, arena + ar->off[SFQ_S_USR_LQLT], ar->cap[SFQ_S_USR_LQLT]);
    x_rec.init(pwpool, SFQ_X_REC, arena + ar->off[SFQ_S_REC_X], ar->cap[SFQ_S_REC_X]);
    x_llen.init(pwpool, SFQ_X_LLEN, arena + ar->off[SFQ_S_USR_X], ar->cap[SFQ_S_USR_X]);
    x_qlen.init(pwpool, SFQ_X_QLEN, arena + ar->off[SFQ_S_USR_XQ], ar->cap[SFQ_S_USR_XQ]);
    x_sgen.init(pwpool, SFQ_X_SGEN, arena + ar->off[SFQ_S_USR_PFG], ar->cap[SFQ_S_USR_PFG]);
    x_sqlt.init(pwpool, SFQ_X_SQLT, arena + ar->off[SFQ_S_USR_PFQ], ar->cap[SFQ_S_USR_PFQ]);

    SfqSpaceMap (&smap)[2] = scr->smap;
    uint8_t (&ctype)[2][65] = scr->ctype;
    uint64_t (&cnumb)[2][65] = scr->cnumb;
    for (int a = 0; a < 2; a++) for (int b = 0; b < 65; b++) { ctype[a][b] = 0; cnumb[a][b] = 0; }
    smap[0].len = smap[1].len = 0;
    uint32_t hflip = 0;                                // which stage holds the current id line
    uint32_t imap = 0;
    uint64_t x_index = 0;                              // RecBase::m_last.index
    const uint32_t solid = meta->solid;
    uint32_t m_llen = (uint32_t)meta->llen;            // UsrSave::m_llen (sticky), usrs.cpp:126-131
    uint64_t i_llen = 0, i_qlen = 0, i_sgen = 0, i_sqlt = 0, i_long = 0;
    uint8_t pf_gen = 0, pf_qlt = 0;
    const uint8_t *prev = nullptr;
    bool have_first = false;                           // RecBase::m_last.initilized
    uint32_t status = SFQ_OK, status_arg = 0;

    for (uint32_t r = 0; r < meta->nrec; r++) {
        const uint64_t recno = (uint64_t)r + 1;        // g_record_count
        const SfqRecView v = sfq_rec_view(text, ls, meta->line0, r, solid);
        if (v.big) {
            // get_oversized_record (usrs.cpp:269-301): the four lines, newlines included, verbatim: id line (without '@')
            // and '+' line to usr.lrec, base line to usr.lgen, quality line to usr.lqlt.  When only the quality line is too
            // long, get_record has already issued this record's prefix / length updates (usrs.cpp:337-364).
            if (v.hlen < SFQ_MAX_ID_LLEN - 1 && v.raw_llen < SFQ_MAX_GN_LLEN - 1) {
                if (solid && pf_gen != v.pf_gen && v.pf_gen) { x_sgen.put(recno - i_sgen); x_sgen.put_chr(v.pf_gen); i_sgen = recno; pf_gen = v.pf_gen; }
                if (m_llen != v.raw_llen) { x_llen.put(recno - i_llen); x_llen.put((uint16_t)v.raw_llen); i_llen = recno; m_llen = (uint16_t)v.raw_llen; }
                if (solid && pf_qlt != v.pf_qlt) { x_sqlt.put(recno - i_sqlt); x_sqlt.put_chr(v.pf_qlt); i_sqlt = recno; pf_qlt = v.pf_qlt; }
            }
            x_lrec.put(recno - i_long); i_long = recno;
            const uint64_t *l = ls + meta->line0 + 4ull * r;
            for (uint64_t k = l[0] + 1; k < l[1]; k++) x_lrec.put_chr(text[k]);
            for (uint64_t k = l[1]; k < l[2]; k++) x_lgen.put_chr(text[k]);
            for (uint64_t k = l[2]; k < l[3]; k++) x_lrec.put_chr(text[k]);
            for (uint64_t k = l[3]; k < l[4]; k++) x_lqlt.put_chr(text[k]);
            continue;
        }
        // ---- framing exceptions, in get_record order (usrs.cpp:320-372)
        if (solid && pf_gen != v.pf_gen && v.pf_gen) {
            x_sgen.put(recno - i_sgen); x_sgen.put_chr(v.pf_gen); i_sgen = recno; pf_gen = v.pf_gen;
        }
        if (m_llen != v.llen) {
            x_llen.put(recno - i_llen); x_llen.put((uint16_t)v.llen); i_llen = recno; m_llen = (uint16_t)v.llen;
        }
        if (solid && pf_qlt != v.pf_qlt) {
            x_sqlt.put(recno - i_sqlt); x_sqlt.put_chr(v.pf_qlt); i_sqlt = recno; pf_qlt = v.pf_qlt;
        }
        if (v.qlen != m_llen) { x_qlen.put(recno - i_qlen); x_qlen.put((uint16_t)v.qlen); i_qlen = recno; }

        // ---- header model (recs.cpp:277-372)
        hflip ^= 1u;
        const uint8_t *buf = sfq_rec_stage(scr->hbuf[hflip], v.hdr, v.hlen);       // (`prev` keeps pointing at the other stage)
        if (!have_first) {
            // first header travels in clear as the `rec.first` info key (recs.cpp:68-75)
            if (v.hlen > 399) { status = SFQ_E_FIRSTHDR; break; }
            have_first = true;
            imap = 0;
            if (!sfq_map_space(smap[0], buf)) { status = SFQ_E_SEPS; status_arg = (uint32_t)recno; break; }
            prev = buf;
            continue;
        }
        const uint32_t pm = imap, im = imap ^ 1u;
        imap = im;
        SfqSpaceMap &S = smap[im];
        const SfqSpaceMap &P = smap[pm];
        if (!sfq_map_space(S, buf)) { status = SFQ_E_SEPS; status_arg = (uint32_t)recno; break; }
        bool same = S.len == P.len;
        for (uint32_t k = 0; same && k < S.len; k++) same = S.str[k] == P.str[k];
        if (!same) {                                                            // recs.cpp:291-304
            x_rec.put(recno - x_index);
            x_index = recno;
            x_rec.put(v.hlen);
            for (uint32_t j = 0; j < v.hlen; j++) x_rec.put_chr(buf[j]);
            for (int b = 0; b < 65; b++) ctype[im][b] = 0;
            prev = buf;
            continue;
        }
        uint64_t map = 0;
        for (uint32_t i = 0; i < S.len; i++) {
            bool diff = S.wln[i] != P.wln[i];
            const uint8_t *a = buf + S.off[i], *b = prev + P.off[i];
            for (uint32_t k = 0; !diff && k < S.wln[i]; k++) diff = a[k] != b[k];
            if (diff) map |= 1ULL << (i & 63u);        // x86 shift-count masking of DO_SET at i == 64
        }
        SfqFieldRangers R0; R0.bind(pwpool, 0);
        R0.num.put(rc, map);                                                    // recs.cpp:312
        for (uint32_t i = 0; i < S.len; i++) {
            if (!(map & (1ULL << (i & 63u)))) {
                ctype[im][i] = ctype[pm][i];
                cnumb[im][i] = cnumb[pm][i];
                continue;
            }
            const uint8_t *b = buf + S.off[i];
            uint64_t bnum;
            uint32_t type = sfq_numberwang(b, S.wln[i], bnum, ctype[pm][i]);
            SfqFieldRangers R; R.bind(pwpool, i + 1);
            if (type == SFQ_ST_STR) {
                R.type.put(rc, type);
                R.num.put(rc, S.wln[i]);
                for (uint32_t j = 0; j < S.wln[i]; j++) R.str.put(rc, b[j]);
                ctype[im][i] = 0;
                continue;
            }
            const uint64_t pnum = ctype[pm][i] ? cnumb[pm][i] : 0;
            uint64_t gap;
            ctype[im][i] = (type < SFQ_ST_STR || type >= SFQ_ST_DGT_Z) ? 1 : 2;
            cnumb[im][i] = bnum;
            if (bnum < pnum) { gap = pnum - bnum; type++; } else gap = bnum - pnum;
            R.type.put(rc, type);
            R.num.put(rc, gap);
        }
        prev = buf;
    }
    rc.finish();
    bool ovf = rc.out.overflow();
    ar->size[SFQ_S_REC] = rc.out.n;
    ar->size[SFQ_S_REC_X] = x_rec.close(ovf);
    ar->size[SFQ_S_USR_X] = x_llen.close(ovf);
    ar->size[SFQ_S_USR_XQ] = x_qlen.close(ovf);
    ar->size[SFQ_S_USR_PFG] = x_sgen.close(ovf);
    ar->size[SFQ_S_USR_PFQ] = x_sqlt.close(ovf);
    ar->size[SFQ_S_USR_LREC] = x_lrec.close(ovf);
    ar->size[SFQ_S_USR_LGEN] = x_lgen.close(ovf);
    ar->size[SFQ_S_USR_LQLT] = x_lqlt.close(ovf);
    if (status == SFQ_OK && ovf) status = SFQ_E_CAP;
    if (status != SFQ_OK && meta->status == SFQ_OK) { meta->status = status; meta->status_arg = status_arg; }
}

// ---------------------------------------------------------------------------- usr decode
// UsrLoad::update (usrs.cpp:471-510): replays usr.x / usr.x.q / usr.pfg / usr.pfq into per-record
// tables so that the qlt, gen and rec decoders of the chunk can run concurrently afterwards.
//
// Oversized records (usr.lrec / usr.lgen / usr.lqlt, usrs.cpp:473-485) are decoded here, before anything else: their
// lines go to the FRONT of the chunk's three planes (id line + '\n' + '+' line as one block in the header plane), their
// table entries carry the lines' lengths with SFQ_BIG_BIT set - the model decoders skip such records, the assemble
// kernel prints them verbatim.  `hdrs` / `bases` / `quals` = the chunk's plane regions, caps = bytes available there.
SFQ_HDN void sfq_usr_decode_chunk(const uint8_t *in, const uint32_t *ssize, const uint64_t *soff,
                                  SfqChunkMeta *meta, uint32_t *pwpool,
                                  uint32_t *llen_tab, uint32_t *qlen_tab, uint8_t *pfg_tab, uint8_t *pfq_tab,
                                  uint32_t *hlen_tab = nullptr, uint64_t *hoff_tab = nullptr, uint64_t *boff_tab = nullptr, uint64_t *qoff_tab = nullptr,
                                  uint8_t *hdrs = nullptr, uint64_t hcap = 0, uint8_t *bases = nullptr, uint64_t bcap = 0,
                                  uint8_t *quals = nullptr, uint64_t qcap = 0) {
    uint32_t status = SFQ_OK;
    uint64_t gh = 0, gb = 0, gq = 0;
    uint32_t nbig = 0;
    if (ssize[SFQ_S_USR_LREC]) {
        if (!hlen_tab) status = SFQ_E_CORRUPT;
        else {
            SfqXLoad x_lrec;
            x_lrec.init(pwpool, SFQ_X_LREC, in + soff[SFQ_S_USR_LREC], ssize[SFQ_S_USR_LREC]);
            uint64_t i_long = x_lrec.get();
            for (uint32_t r = 0; r < meta->nrec; r++) hlen_tab[r] = 0;
            while (i_long && status == SFQ_OK) {
                if (i_long > meta->nrec) { status = SFQ_E_CORRUPT; break; }
                const uint32_t r = (uint32_t)(i_long - 1);
                const uint64_t start = gh;
                for (int line = 0; line < 2 && status == SFQ_OK; line++) {          // id line, '+' line
                    for (;;) {
                        const uint8_t c = x_lrec.get_chr();
                        if (c == '\n' && line == 1) break;
                        if (gh >= hcap) { status = SFQ_E_CORRUPT; break; }
                        hdrs[gh++] = c;
                        if (c == '\n') break;
                    }
                }
                hlen_tab[r] = (uint32_t)(gh - start) | SFQ_BIG_BIT;
                hoff_tab[r] = start;
                nbig++;
                const uint64_t d = x_lrec.get();
                if (!d) break;
                i_long += d;
            }
            SfqXLoad x_lgen, x_lqlt;
            x_lgen.init(pwpool, SFQ_X_LGEN, in + soff[SFQ_S_USR_LGEN], ssize[SFQ_S_USR_LGEN]);
            x_lqlt.init(pwpool, SFQ_X_LQLT, in + soff[SFQ_S_USR_LQLT], ssize[SFQ_S_USR_LQLT]);
            if (nbig && (!x_lgen.rc.valid || !x_lqlt.rc.valid)) status = SFQ_E_CORRUPT;
            for (uint32_t r = 0; r < meta->nrec && status == SFQ_OK; r++) {
                if (!(hlen_tab[r] & SFQ_BIG_BIT)) continue;
                uint64_t start = gb;
                for (;;) { const uint8_t c = x_lgen.get_chr(); if (c == '\n') break; if (gb >= bcap) { status = SFQ_E_CORRUPT; break; } bases[gb++] = c; }
                llen_tab[r] = (uint32_t)(gb - start) | SFQ_BIG_BIT; boff_tab[r] = start;
                start = gq;
                for (;;) { const uint8_t c = x_lqlt.get_chr(); if (c == '\n') break; if (gq >= qcap) { status = SFQ_E_CORRUPT; break; } quals[gq++] = c; }
                qlen_tab[r] = (uint32_t)(gq - start) | SFQ_BIG_BIT; qoff_tab[r] = start;
            }
        }
    }
    SfqXLoad x_llen, x_qlen, x_sgen, x_sqlt;
    x_llen.init(pwpool, SFQ_X_LLEN, in + soff[SFQ_S_USR_X], ssize[SFQ_S_USR_X]);
    x_qlen.init(pwpool, SFQ_X_QLEN, in + soff[SFQ_S_USR_XQ], ssize[SFQ_S_USR_XQ]);
    x_sgen.init(pwpool, SFQ_X_SGEN, in + soff[SFQ_S_USR_PFG], ssize[SFQ_S_USR_PFG]);
    x_sqlt.init(pwpool, SFQ_X_SQLT, in + soff[SFQ_S_USR_PFQ], ssize[SFQ_S_USR_PFQ]);
    const bool solid = meta->solid != 0;
    uint64_t m_llen = (uint64_t)(uint32_t)meta->llen, m_qlen = m_llen;
    uint64_t i_llen = x_llen.get(), i_qlen = x_qlen.get(), i_sgen = x_sgen.get(), i_sqlt = x_sqlt.get();
    uint8_t pf_gen = 0, pf_qlt = 0;
    uint64_t nb = 0, nq = 0;
    for (uint32_t r = 0; r < meta->nrec && status == SFQ_OK; r++) {
        const uint64_t recno = (uint64_t)r + 1;
        if (nbig && (hlen_tab[r] & SFQ_BIG_BIT)) { pfg_tab[r] = 0; pfq_tab[r] = 0; continue; }   // (update() starts over at the next record number)
        if (hlen_tab) hlen_tab[r] = 0;
        if (i_llen == recno) { m_llen = x_llen.get(); m_qlen = m_llen; i_llen += x_llen.get(); }
        if (i_qlen == recno) { m_qlen = x_qlen.get(); i_qlen += x_qlen.get(); }
        else if (m_qlen != m_llen) m_qlen = m_llen;
        if (solid && i_sgen == recno) { pf_gen = x_sgen.get_chr(); i_sgen += x_sgen.get(); }
        if (solid && i_sqlt == recno) { pf_qlt = x_sqlt.get_chr(); i_sqlt += x_sqlt.get(); }
        if (m_llen >= SFQ_MAX_GN_LLEN || m_qlen >= SFQ_MAX_GN_LLEN) { status = SFQ_E_CORRUPT; m_llen = m_qlen = 0; }
        llen_tab[r] = (uint32_t)m_llen;
        qlen_tab[r] = (uint32_t)m_qlen;
        pfg_tab[r] = pf_gen;
        pfq_tab[r] = pf_qlt;
        nb += m_llen; nq += m_qlen;
    }
    if (meta->pad & 1u) {                     // imported from a reference file: the header holds upper bounds
        if (nb > meta->nbases || nq > meta->nquals) status = SFQ_E_CORRUPT;
        else { meta->nbases = (uint32_t)nb; meta->nquals = (uint32_t)nq; }
    } else if (nb != meta->nbases || nq != meta->nquals || nbig != meta->nbig) status = SFQ_E_CORRUPT;
    meta->nbig = nbig; meta->big_bases = (uint32_t)gb; meta->big_quals = (uint32_t)gq; meta->big_hdr = (uint32_t)gh;
    if (status != SFQ_OK && meta->status == SFQ_OK) meta->status = status;
}

// ---------------------------------------------------------------------------- rec decode
SFQ_HD uint32_t sfq_fmt_u64(uint8_t *b, uint64_t v, uint32_t base, bool upper, bool is_signed) {
    // what sprintf("%lld" / "%llx" / "%llX") prints for a non-zero value (recs.cpp:436-456)
    uint8_t tmp[24];
    uint32_t n = 0, o = 0;
    if (is_signed && (int64_t)v < 0) { b[o++] = '-'; v = 0ull - v; }
    while (v) {
        const uint32_t d = (uint32_t)(v % base);
        tmp[n++] = (uint8_t)(d < 10 ? '0' + d : (upper ? 'A' : 'a') + d - 10);
        v /= base;
    }
    while (n) b[o++] = tmp[--n];
    return o;
}

// Decodes all headers of a chunk into `hdrs` (each followed by '\n'; record r at hoff_tab[r],
// length hlen_tab[r]).  `hcap` = bytes available in the plane; the caller gives SFQ_HDR_PLANE(meta)
// so the conservative room checks below never reject a valid container.
SFQ_HDN void sfq_rec_decode_chunk(const uint8_t *in, const uint32_t *ssize, const uint64_t *soff,
                                  SfqChunkMeta *meta, uint32_t *pwpool, const uint8_t *rec_first,
                                  uint32_t rec_first_len, uint8_t *hdrs, uint64_t hcap,
                                  uint32_t *hlen_tab, uint64_t *hoff_tab) {
    SfqDec rc;
    rc.start(in + soff[SFQ_S_REC], ssize[SFQ_S_REC]);
    SfqXLoad x_rec;
    x_rec.init(pwpool, SFQ_X_REC, in + soff[SFQ_S_REC_X], ssize[SFQ_S_REC_X]);
    SfqSpaceMap S;
    uint8_t ctype[2][65];
    uint64_t cnumb[2][65];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 65; b++) { ctype[a][b] = 0; cnumb[a][b] = 0; }
    uint32_t imap = 0;
    uint64_t x_index = x_rec.get();                                             // recs.cpp:104-105
    uint64_t pos = meta->nbig ? meta->big_hdr : 0;                              // oversized records' lines sit at the front of the plane
    const bool pre5 = (meta->pad & 2u) != 0;                                    // SFQ_BLOB_PRE5
    const uint8_t *prev = nullptr;
    bool have_first = false;
    uint32_t status = SFQ_OK;

    for (uint32_t r = 0; r < meta->nrec; r++) {
        const uint64_t recno = (uint64_t)r + 1;
        if (meta->nbig && (hlen_tab[r] & SFQ_BIG_BIT)) continue;                // (UsrLoad::update prints it and moves on, usrs.cpp:473-485)
        uint8_t *buf = hdrs + pos;
        // every branch below checks room before it writes; a valid container never trips it
        uint64_t room = hcap - pos;
        uint32_t n = 0;
        if (!have_first) {                                                      // recs.cpp:375-381
            have_first = true;
            if (rec_first_len + 1ull > room) { status = SFQ_E_CORRUPT; break; }
            for (uint32_t k = 0; k < rec_first_len; k++) buf[k] = rec_first[k];
            n = rec_first_len;
            imap = 0;
        } else {
            const uint32_t pm = imap, im = imap ^ 1u;
            imap = im;
            if (x_index == recno) {                                             // recs.cpp:386-393
                const uint64_t len = x_rec.get();
                if (len + 1ull > room) { status = SFQ_E_CORRUPT; break; }
                for (uint32_t j = 0; j < (uint32_t)len; j++) buf[j] = x_rec.get_chr();
                x_index += x_rec.get();
                for (int b = 0; b < 65; b++) ctype[im][b] = 0;
                n = (uint32_t)len;
            } else {
                if (!sfq_map_space(S, prev)) { status = SFQ_E_CORRUPT; break; }
                SfqFieldRangers R0; R0.bind(pwpool, 0);
                const uint64_t map = R0.num.get(rc);
                uint8_t *b = buf;
                bool bad = false;
                for (uint32_t i = 0; i < S.len; i++) {
                    if ((uint64_t)(b - buf) + S.wln[i] + 44ull > room) { bad = true; break; }
                    if (pre5) {                                                 // RecLoad::load_pre5, recs.cpp:463-510
                        if (map & (1ULL << (i & 63u))) {
                            SfqFieldRangers R; R.bind(pwpool, i + 1);
                            const uint32_t type = R.type.get(rc);
                            if (type == SFQ_ST_DGT || type == SFQ_ST_DLT) {
                                // the previous header's field, read as a number (is_number, recs.cpp:265-275: no leading zero)
                                const uint8_t *pf = prev + S.off[i];
                                uint64_t pval = 0;
                                bool isnum = pf[0] != '0';
                                for (uint32_t k = 0; k < S.wln[i] && isnum; k++) { if (sfq_is_dig(pf[k])) pval = pval * 10u + (uint64_t)(pf[k] - '0'); else isnum = false; }
                                if (!isnum) { bad = true; break; }              // (the reference asserts)
                                const uint64_t gap = R.num.get(rc);
                                const uint64_t val = type == SFQ_ST_DGT ? pval + gap : pval - gap;
                                if (val == 0) *b++ = '0'; else b += sfq_fmt_u64(b, val, 10u, false, true);
                            } else if (type == SFQ_ST_STR) {
                                const uint64_t len = R.num.get(rc);
                                if ((uint64_t)(b - buf) + len + 2ull > room) { bad = true; break; }
                                for (uint32_t j = 0; j < (uint32_t)len; j++) *b++ = (uint8_t)R.str.get(rc);
                            } else { bad = true; break; }
                        } else {
                            const uint8_t *src = prev + S.off[i];
                            for (uint32_t k = 0; k < S.wln[i]; k++) *b++ = src[k];
                        }
                        *b++ = S.str[i];
                        continue;
                    }
                    if (!(map & (1ULL << (i & 63u)))) {
                        const uint8_t *src = prev + S.off[i];
                        for (uint32_t k = 0; k < S.wln[i]; k++) *b++ = src[k];
                        *b++ = S.str[i];
                        ctype[im][i] = ctype[pm][i];
                        cnumb[im][i] = cnumb[pm][i];
                        continue;
                    }
                    SfqFieldRangers R; R.bind(pwpool, i + 1);
                    const uint32_t type = R.type.get(rc);
                    if (type == SFQ_ST_STR) {
                        const uint64_t len = R.num.get(rc);
                        if ((uint64_t)(b - buf) + len + 2ull > room) { bad = true; break; }
                        for (uint32_t j = 0; j < (uint32_t)len; j++) *b++ = (uint8_t)R.str.get(rc);
                        ctype[im][i] = 0;
                        *b++ = S.str[i];
                        continue;
                    }
                    if (type > SFQ_ST_DLT_Z) { bad = true; break; }
                    const uint64_t pval = ctype[pm][i] == 0 ? 0 : cnumb[pm][i];
                    const uint64_t gap = R.num.get(rc);
                    const bool less = type == SFQ_ST_DLT || type == SFQ_ST_HLT || type == SFQ_ST_HLT_Z ||
                                      type == SFQ_ST_HLTC || type == SFQ_ST_HLTC_Z || type == SFQ_ST_DLT_Z;
                    const uint64_t val = less ? pval - gap : pval + gap;
                    ctype[im][i] = (type < SFQ_ST_STR || type >= SFQ_ST_DGT_Z) ? 1 : 2;
                    cnumb[im][i] = val;
                    if (val == 0) *b++ = '0';                                   // recs.cpp:453-454
                    else {
                        const bool dec = type <= SFQ_ST_DLT || type >= SFQ_ST_DGT_Z;
                        const bool zed = type == SFQ_ST_HGT_Z || type == SFQ_ST_HLT_Z || type == SFQ_ST_HGTC_Z ||
                                         type == SFQ_ST_HLTC_Z || type == SFQ_ST_DGT_Z || type == SFQ_ST_DLT_Z;
                        const bool upper = type >= SFQ_ST_HGTC && type <= SFQ_ST_HLTC_Z;
                        if (zed) *b++ = '0';
                        b += sfq_fmt_u64(b, val, dec ? 10u : 16u, upper, dec);
                    }
                    *b++ = S.str[i];
                }
                if (bad) { status = SFQ_E_CORRUPT; break; }
                n = (uint32_t)(b - buf) - 1u;                                   // recs.cpp:460
            }
        }
        buf[n] = '\n';       // terminator the next record's tokeniser stops at (UsrLoad::putline)
        hlen_tab[r] = n;
        hoff_tab[r] = pos;
        prev = buf;
        pos += (uint64_t)n + 1;
    }
    // pos may differ from hdr_bytes + nrec only for headers the reference itself does not reproduce
    // (e.g. a decimal field >= 2^63 comes back through "%lld" with a sign, recs.cpp:436-456)
    if (status != SFQ_OK && meta->status == SFQ_OK) meta->status = status;
}
