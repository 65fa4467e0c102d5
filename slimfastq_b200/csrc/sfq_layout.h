// Host-side sizing policy for the per-chunk device workspace (shared by the library and the test
// emulation so both exercise the same table geometry).
#pragma once
#include <stdint.h>
#include "sfq_common.cuh"

static inline uint32_t sfq_ceil_log2(uint64_t v) { uint32_t b = 0; while ((1ull << b) < v) b++; return b; }

// Quality-context table: 4096 contexts at level 1, 65536 otherwise (qlts.cpp:34-39), 256 B each.
static inline uint64_t sfq_qtable_bytes(int level) { return (level <= 1 ? 4096ull : 65536ull) * SFQ_L64_WORDS * 4; }
// Hashed quality-context table of the lane-cooperative coder: a 1 MiB chunk touches ~4-8 k of the
// 65536 contexts at level 3/4 and 10-27 k at level 2, so 2^14 / 2^15 entries to start with; a chunk
// that fills its table is rerun with one more bit (16 bits = the direct table).
static inline uint32_t sfq_q_cbits(int level, uint32_t grow) {
    if (level <= 1) return 12;
    uint32_t b = (level == 2 ? 15u : 14u) + grow;
    return b > 16 ? 16 : b;
}
// Decoder-side quality-context hash: entries for a load factor <= 1/2 at `contexts` visited contexts
// (recorded in the blob header; every extra probe is a memory round trip on the decoder's serial chain),
// or the direct table when that would be as large.
static inline uint32_t sfq_q_entries(int level, uint64_t contexts, uint32_t grow) {
    if (level <= 1) return 4096;
    uint64_t e = (2 * contexts + 64) << grow;
    return (uint32_t)(e >= 65536 ? 65536 : e);
}
static inline uint64_t sfq_qhash_bytes(int level, uint32_t cbits) { return (level <= 1 ? 4096ull : (1ull << cbits)) * SFQ_L64_WORDS * 4; }
// Base-context table.  Level 1: direct 2^18 x u32.  Levels 2-4: open-addressing hash of 64-bit
// slots sized for a load factor <= 0.5 at `max_bases` insertions, never larger than the dense table
// of the level would be.
static inline uint32_t sfq_gen_hbits(int level, uint64_t max_bases, uint32_t grow) {
    if (level <= 1) return 18;
    uint32_t b = sfq_ceil_log2(max_bases * 2 + 16) + grow;
    return b < 10 ? 10 : b > 27 ? 27 : b;
}
// Two-phase encoder: partitions of the base-context space per chunk, for at most 1024 bases per partition on average
// (k_gen_replay keeps a partition's contexts in a 2048-slot table in shared memory); a table that still fills up reruns
// with twice as many.
static inline uint32_t sfq_gen_gp_bits(uint64_t max_bases, uint32_t grow) {
    uint32_t b = sfq_ceil_log2((max_bases + 1023) / 1024) + grow;
    return b > 14 ? 14 : b;
}
static inline uint64_t sfq_gtable_bytes(int level, uint32_t hbits) {
    return level <= 1 ? (1ull << 18) * 4 : (1ull << hbits) * 8;
}
// Decoder-side base-context table (SfqGenBuckets): buckets of four 64-bit slots, sized for a load
// factor <= 2/3 at `contexts` distinct contexts (the blob header records how many a chunk touched).
static inline uint32_t sfq_gen_nbuckets(int level, uint64_t contexts, uint32_t grow) {
    if (level <= 1) return 1;
    uint64_t slots = (contexts + (contexts >> 1) + 64) << grow;
    uint64_t nb = ((slots + 15) / 16) * 4;                      // whole 128-byte lines of four buckets
    return (uint32_t)(nb > 0x7ffffffcull ? 0x7ffffffcull : nb);
}
static inline uint64_t sfq_gbuckets_bytes(int level, uint32_t nbuckets) {
    return level <= 1 ? (1ull << 18) * 4 : (uint64_t)nbuckets * 32;
}
static inline uint64_t sfq_pwpool_bytes() { return (uint64_t)SFQ_PW_PER_CHUNK * SFQ_PW_WORDS * 4; }

// Output arena capacities.  Generous versus the typical 2 bit/base and 2-5 bit/quality; a stream
// that still overflows is reported (SFQ_E_CAP) and the host retries the wave with `grow` doubled.
static inline void sfq_arena_layout(const SfqChunkMeta *m, uint32_t grow, uint64_t base, SfqArena *a, uint64_t *end) {
    const uint64_t g = 1ull << grow;
    uint64_t cap[SFQ_NSTREAMS];
    // first try: what real data needs with room to spare (2.25 bit per base, 6 bit per quality, half of the header bytes, one
    // N exception per 64 bases, a length exception for every record); every retry doubles
    cap[SFQ_S_REC] = ((uint64_t)m->hdr_bytes * g) / 2 + 4ull * m->nrec * g + 256;
    cap[SFQ_S_GEN] = ((uint64_t)m->nbases * g * 9) / 32 + 256;
    cap[SFQ_S_QLT] = ((uint64_t)m->nquals * g * 3) / 4 + 256;
    cap[SFQ_S_GEN_NS] = ((uint64_t)m->nbases * g) / 64 + 256;
    cap[SFQ_S_GEN_NN] = ((uint64_t)m->nbases * g) / 64 + 256;
    cap[SFQ_S_REC_X] = ((uint64_t)m->hdr_bytes * g) / 4 + 4ull * m->nrec * g + 256;
    cap[SFQ_S_USR_X] = 4ull * m->nrec * g + 256;
    cap[SFQ_S_USR_XQ] = 4ull * m->nrec * g + 256;
    cap[SFQ_S_USR_PFG] = 2ull * m->nrec * g + 256;
    cap[SFQ_S_USR_PFQ] = 2ull * m->nrec * g + 256;
    // oversized records, character by character through an adaptive 256-symbol model: rarely more than a byte per character
    cap[SFQ_S_USR_LREC] = m->nbig ? ((uint64_t)m->big_hdr + 16ull * m->nbig) * g * 2 + 256 : 64;
    cap[SFQ_S_USR_LGEN] = m->nbig ? ((uint64_t)m->big_bases + 2ull * m->nbig) * g * 2 + 256 : 64;
    cap[SFQ_S_USR_LQLT] = m->nbig ? ((uint64_t)m->big_quals + 2ull * m->nbig) * g * 2 + 256 : 64;
    uint64_t o = base;
    for (int k = 0; k < SFQ_NSTREAMS; k++) {
        a->off[k] = o;
        a->cap[k] = (uint32_t)(cap[k] > 0xFFFFFF00ull ? 0xFFFFFF00ull : cap[k]);
        a->size[k] = 0;
        o += (a->cap[k] + 15ull) & ~15ull;
    }
    *end = o;
}
