// The reference's own .sfq file: an 8 KiB-page write-once container (filer.cpp:41-303, filer.hpp:34-41).
// Host-only format code (no coding happens here): lets a single-chunk container of this library be
// written as a file the unmodified reference binary decodes, and a reference-written file be read
// into a single-chunk container the GPU path decodes.
//
//   page 0   first page of the info stream ("key=value\n" lines, first line whoami=slimfastq)
//   page 1   file table: 341 packed entries {u64 name, u64 size, u32 first, u32 node}; entry 0 is the
//            info stream and its `first` field holds the entry count on disk (filer.cpp:104-105,130)
//   a stream = its `first` page, then the data pages listed in its node page(s); a node page is
//            u32[2048]: 2047 data pages + the next node page (filer.cpp:210-242, 273-303)
// The loader only follows these pointers, so the writer below lays every stream out contiguously.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "sfq_container.h"

#define SFQ_WORM_PAGE   0x2000u
#define SFQ_WORM_NODES  2047u             // data-page slots of a node page
#define SFQ_WORM_FILES  341u              // entries of the table page

#pragma pack(push, 1)
struct SfqWormEntry { uint64_t name, size; uint32_t first, node; };
#pragma pack(pop)

static const char *const kSfqStreamNames[SFQ_NSTREAMS] = {"rec", "gen", "qlt", "gen.Ns", "gen.Nn", "rec.x", "usr.x", "usr.x.q", "usr.pfg", "usr.pfq",
                                                          "usr.lrec", "usr.lgen", "usr.lqlt"};

// pages a stream of n bytes occupies: data pages + node pages
static inline uint64_t sfq_worm_pages(uint64_t n) {
    const uint64_t data = n ? (n + SFQ_WORM_PAGE - 1) / SFQ_WORM_PAGE : 1;
    const uint64_t listed = data - 1;
    return data + (listed + SFQ_WORM_NODES - 1) / SFQ_WORM_NODES;
}

// Appends one stream at page `*next`; fills its table entry.
static inline void sfq_worm_put_stream(std::vector<uint8_t> &file, uint64_t *next, const uint8_t *p, uint64_t n,
                                       uint32_t first_page, SfqWormEntry *e) {
    auto page_at = [&](uint64_t pg) -> uint8_t * {
        if (file.size() < (pg + 1) * SFQ_WORM_PAGE) file.resize((pg + 1) * SFQ_WORM_PAGE, 0);
        return file.data() + pg * SFQ_WORM_PAGE;
    };
    e->size = n; e->first = first_page; e->node = 0;
    const uint64_t first_n = n < SFQ_WORM_PAGE ? n : SFQ_WORM_PAGE;
    if (first_n) memcpy(page_at(first_page), p, first_n); else page_at(first_page);
    uint64_t done = first_n;
    uint64_t node_pg = 0;
    uint32_t slot = 0;
    while (done < n) {
        if (!node_pg || slot == SFQ_WORM_NODES) {               // a new node page
            const uint64_t nn = (*next)++;
            page_at(nn);
            if (!node_pg) e->node = (uint32_t)nn;
            else { const uint32_t v = (uint32_t)nn; memcpy(page_at(node_pg) + 4 * SFQ_WORM_NODES, &v, 4); }
            node_pg = nn; slot = 0;
        }
        const uint64_t dp = (*next)++;
        const uint64_t k = n - done < SFQ_WORM_PAGE ? n - done : SFQ_WORM_PAGE;
        memcpy(page_at(dp), p + done, k);
        const uint32_t v = (uint32_t)dp;
        memcpy(page_at(node_pg) + 4 * slot, &v, 4);
        slot++;
        done += k;
    }
}

// One chunk (blob header, rec.first, stream pointers) -> reference file bytes.
static inline bool sfq_worm_write(const SfqBlobHeader &b, const uint8_t *rec_first, const uint8_t *const stream[SFQ_NSTREAMS],
                                  const char *orig_filename, std::vector<uint8_t> &file, std::string &err) {
    if (b.rec_first_len >= 0x200) { err = "oversize string value"; return false; }      // config.cpp:134-140
    std::string info;
    auto add = [&](const char *k, const std::string &v) { info += k; info += '='; info += v; info += '\n'; };
    add("whoami", "slimfastq");
    add("version", std::to_string(SFQ_INTERNAL_VERSION));
    add("config.level", std::to_string(b.level));
    {   // the reference reads an info line into 512 bytes and refuses values of 0x1ff+ characters (config.cpp:87-107,134-140):
        // a longer path travels as its last component, cut if need be
        std::string fn = orig_filename && *orig_filename ? orig_filename : "<< stdin >>";
        if (fn.size() >= 0x1e0) { const size_t sl = fn.rfind('/'); if (sl != std::string::npos) fn = fn.substr(sl + 1); }
        if (fn.size() >= 0x1e0) fn.resize(0x1df);
        add("orig.filename", fn);
    }
    add("orig.size", std::to_string((unsigned long long)b.text_len));
    if (b.solid) add("usr.solid", "1");
    if (b.nbig < b.nrec || b.llen) {               // (a file of oversized records only has neither key, usrs.cpp:190-198)
        add("llen", std::to_string(b.llen));
        add("usr.2id", b.two_id ? "1" : "0");
    }
    if (b.nbig < b.nrec) add("rec.first", std::string((const char *)rec_first, b.rec_first_len));
    if (b.n_byte) add("gen.N_byte", std::to_string((unsigned)b.n_byte));
    add("num_records", std::to_string(b.nrec));
    uint64_t pages = 0;
    std::string full;
    for (int it = 0; it < 3; it++) {          // comp.size counts the info stream's own pages
        full = info + "comp.size=" + std::to_string((unsigned long long)(pages * SFQ_WORM_PAGE)) + "\n";
        uint64_t p = 1 + sfq_worm_pages(full.size());           // table page + info stream
        for (int k = 0; k < SFQ_NSTREAMS; k++) if (b.ssize[k]) p += sfq_worm_pages(b.ssize[k]);
        if (p == pages) break;
        pages = p;
    }
    file.assign(2 * SFQ_WORM_PAGE, 0);
    std::vector<SfqWormEntry> tab(SFQ_WORM_FILES + 1);
    memset(tab.data(), 0, tab.size() * sizeof(SfqWormEntry));
    uint64_t next = 2;
    sfq_worm_put_stream(file, &next, (const uint8_t *)full.data(), full.size(), 0, &tab[0]);
    uint32_t count = 1;
    for (int k = 0; k < SFQ_NSTREAMS; k++) {
        if (!b.ssize[k]) continue;
        SfqWormEntry &e = tab[count++];
        char nm[8]; memset(nm, 0, 8); strncpy(nm, kSfqStreamNames[k], 8);
        memcpy(&e.name, nm, 8);
        const uint32_t first = (uint32_t)next++;
        sfq_worm_put_stream(file, &next, stream[k], b.ssize[k], first, &e);
    }
    tab[0].first = count;                                       // filer.cpp:130
    if (file.size() < next * SFQ_WORM_PAGE) file.resize(next * SFQ_WORM_PAGE, 0);
    memcpy(file.data() + SFQ_WORM_PAGE, tab.data(), SFQ_WORM_PAGE);
    if (next != pages) { err = "internal error: page count mismatch"; return false; }
    return true;
}

static inline bool sfq_worm_get_stream(const uint8_t *f, size_t n, const SfqWormEntry &e, uint32_t first, std::vector<uint8_t> &out) {
    out.clear();
    if (e.size == 0) return true;
    if (e.size > n) return false;
    auto page = [&](uint64_t pg) -> const uint8_t * { return (pg + 1) * SFQ_WORM_PAGE <= n ? f + pg * SFQ_WORM_PAGE : nullptr; };
    out.reserve(e.size);
    const uint8_t *p = page(first);
    if (!p) return false;
    uint64_t k = e.size < SFQ_WORM_PAGE ? e.size : SFQ_WORM_PAGE;
    out.insert(out.end(), p, p + k);
    uint64_t node = e.node;
    uint64_t hops = 0;
    while (out.size() < e.size) {
        const uint8_t *np = node ? page(node) : nullptr;
        if (!np || ++hops > n / SFQ_WORM_PAGE) return false;
        for (uint32_t s = 0; s < SFQ_WORM_NODES && out.size() < e.size; s++) {
            uint32_t dp; memcpy(&dp, np + 4 * s, 4);
            const uint8_t *d = page(dp);
            if (!d) return false;
            k = e.size - out.size() < SFQ_WORM_PAGE ? e.size - out.size() : SFQ_WORM_PAGE;
            out.insert(out.end(), d, d + k);
        }
        uint32_t nx; memcpy(&nx, np + 4 * SFQ_WORM_NODES, 4);
        node = nx;
    }
    return true;
}

// Reference file -> info map (first value of a duplicated key wins, config.cpp:100-107) + named streams.
static inline bool sfq_worm_read(const uint8_t *f, size_t n, std::map<std::string, std::string> &info,
                                 std::map<std::string, std::vector<uint8_t>> &streams, std::string &err) {
    if (n < 2 * SFQ_WORM_PAGE || memcmp(f, SFQ_STAMP, 16)) { err = "not a slimfastq file"; return false; }
    SfqWormEntry tab[SFQ_WORM_FILES];
    memcpy(tab, f + SFQ_WORM_PAGE, sizeof tab);
    const uint32_t count = tab[0].first;
    if (count < 1 || count > SFQ_WORM_FILES) { err = "corrupt file table"; return false; }
    std::vector<uint8_t> raw;
    if (!sfq_worm_get_stream(f, n, tab[0], 0, raw)) { err = "corrupt info stream"; return false; }
    size_t pos = 0;
    while (pos < raw.size()) {
        size_t e = pos;
        while (e < raw.size() && raw[e] != '\n') e++;
        const std::string line((const char *)raw.data() + pos, e - pos);
        const size_t eq = line.find('=');
        if (eq != std::string::npos) info.insert(std::make_pair(line.substr(0, eq), line.substr(eq + 1)));
        pos = e + 1;
    }
    for (uint32_t i = 1; i < count; i++) {
        char nm[9]; memcpy(nm, &tab[i].name, 8); nm[8] = 0;
        std::vector<uint8_t> s;
        if (!sfq_worm_get_stream(f, n, tab[i], tab[i].first, s)) { err = std::string("corrupt stream ") + nm; return false; }
        streams[nm] = std::move(s);
    }
    return true;
}
