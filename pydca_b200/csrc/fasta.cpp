// fasta.cpp -- FASTA ingest, integer encoding and first-occurrence de-duplication.
//
// Replaces  PlmDCA::readSequencesFromFile   pydca/plmdca/plmdca_numerics.cpp:685-767
//           alignment_letter2int            pydca/fasta_reader/fasta_reader.py:122-163
//
// Semantics kept: letters are upper-cased (:752 / fasta_reader.py:141), encoded
// with the C++ tables (protein ACDEFGHIKLMNPQRSTVWY -> 0..19, RNA ACGU -> 0..3,
// everything else -> gap = q-1), and only the first occurrence of each distinct
// sequence is kept, in file order (:756 / fasta_reader.py:153).
// Deliberate deltas ("reader semantics"):
//   * records may span several lines (Biopython semantics); the C++ reader
//     would treat every line as a sequence.  Identical for unwrapped files.
//   * letters missing from the C++ map (e.g. 'T' in RNA) are gaps, as in the
//     Python reader; the C++ reader aborts on them (:752).
//   * de-duplication is hash-then-verify, O(N L) instead of O(N^2 L).
#include "dca_common.cuh"
#include "dca_internal.h"

#include <cctype>
#include <cstdio>
#include <string>
#include <cstring>
#include <vector>

namespace dca {

int encode_residue(int biomolecule, int c)
{
    static const char prot[] = "ACDEFGHIKLMNPQRSTVWY";
    static const char rna[] = "ACGU";
    const char *alpha = (biomolecule == 1) ? prot : rna;
    const int q = (biomolecule == 1) ? 21 : 5;
    const int u = std::toupper(c);
    if (u == 0) return q - 1;
    const char *p = std::strchr(alpha, u);
    return p ? (int)(p - alpha) : q - 1;
}

// 64-bit hash of an encoded record, 8 bytes per step (only used to find candidate duplicates:
// equality is always confirmed with memcmp)
static uint64_t record_hash(const uint8_t *p, size_t n)
{
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)n;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    uint64_t w = 0;
    if (i < n) std::memcpy(&w, p + i, n - i);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    return h ^ (h >> 29);
}

long read_fasta_unique(const char *path, int biomolecule, int L, std::vector<uint8_t> &out)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) { set_error("Unable to open file %s", path); return -1; }
    // one read of the whole file, then a single pass over the bytes (the first version went
    // through fgets + std::string per line: 2.3 ms of the 44 ms end-to-end call at N = 5000)
    std::vector<char> text;
    {
        std::fseek(f, 0, SEEK_END);
        const long size = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (size > 0) {
            text.resize((size_t)size);
            const size_t got = std::fread(text.data(), 1, (size_t)size, f);
            text.resize(got);
        } else {                                   // not seekable (pipe): read in pieces
            char buf[1 << 16];
            size_t got;
            while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.insert(text.end(), buf, buf + got);
        }
        std::fclose(f);
    }
    uint8_t lut[256];                              // 255 = white space inside a line (skipped)
    for (int c = 0; c < 256; ++c) lut[c] = (uint8_t)encode_residue(biomolecule, c);
    lut[(unsigned char)'\r'] = lut[(unsigned char)' '] = lut[(unsigned char)'\t'] = 255;

    out.clear();
    out.reserve(text.size());
    // open-addressing table of record indices keyed by the FNV-1a hash of the encoded record
    size_t cap = 1024;
    while (cap < text.size() / (size_t)(L > 0 ? L : 1) * 2 + 16) cap <<= 1;
    std::vector<long> slot(cap, -1);
    std::vector<uint64_t> slot_hash(cap, 0);
    std::vector<uint8_t> rec((size_t)L);
    long n = 0;
    int fill = -1;                 // -1: no record open yet
    long line_no = 0;
    bool bad = false;

    auto flush = [&]() {
        if (fill <= 0) return;     // empty record: skipped like the reference's empty lines
        if (fill < L) { set_error("%s: record ending at line %ld has %d residues, expected %d", path, line_no, fill, L); bad = true; return; }
        const uint64_t h = record_hash(rec.data(), (size_t)L);
        size_t k = (size_t)h & (cap - 1);
        while (slot[k] >= 0) {                     // first occurrence wins (plmdca_numerics.cpp:756)
            if (slot_hash[k] == h && std::memcmp(out.data() + (size_t)slot[k] * L, rec.data(), (size_t)L) == 0) return;
            k = (k + 1) & (cap - 1);
        }
        slot[k] = n; slot_hash[k] = h;
        out.insert(out.end(), rec.begin(), rec.end());
        ++n;
    };

    const char *p = text.data(), *end = p + text.size();
    while (p < end && !bad) {
        const char *eol = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        const char *stop = eol ? eol : end;
        ++line_no;
        if (p < stop && *p == '>') { flush(); fill = 0; }
        else {
            if (fill < 0 && p < stop) fill = 0;     // headerless file: one record
            const size_t len = (size_t)(stop - p);
            // lines without blanks (every line of a normal alignment) are translated in one
            // branch-free run; the byte-by-byte loop below was 60 % of the reader's time
            if (!std::memchr(p, '\r', len) && !std::memchr(p, ' ', len) && !std::memchr(p, '\t', len)) {
                const size_t room = (fill < L) ? (size_t)(L - fill) : 0, take = len < room ? len : room;
                uint8_t *dst = rec.data() + fill;
                const unsigned char *src = reinterpret_cast<const unsigned char *>(p);
                for (size_t k = 0; k < take; ++k) dst[k] = lut[src[k]];
                fill += (int)len;
            } else {
                for (const char *c = p; c < stop; ++c) {
                    const uint8_t v = lut[(unsigned char)*c];
                    if (v == 255) continue;
                    if (fill < L) rec[(size_t)fill] = v;
                    ++fill;
                }
            }
        }
        p = eol ? eol + 1 : end;
    }
    if (!bad) flush();
    if (bad) return -2;
    if (n == 0) { set_error("%s: no sequences found", path); return -3; }
    return n;
}

}  // namespace dca
