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
#include <unordered_map>
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

static uint64_t fnv1a(const uint8_t *p, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

long read_fasta_unique(const char *path, int biomolecule, int L, std::vector<uint8_t> &out)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) { set_error("Unable to open file %s", path); return -1; }
    uint8_t lut[256];
    for (int c = 0; c < 256; ++c) lut[c] = (uint8_t)encode_residue(biomolecule, c);

    out.clear();
    std::unordered_multimap<uint64_t, long> seen;
    std::vector<uint8_t> rec((size_t)L);
    long n = 0;
    int fill = -1;                 // -1: no record open yet
    long line_no = 0;
    bool bad = false;

    auto flush = [&]() {
        if (fill <= 0) return;     // empty record: skipped like the reference's empty lines
        if (fill < L) { set_error("%s: record ending at line %ld has %d residues, expected %d", path, line_no, fill, L); bad = true; return; }
        const uint64_t h = fnv1a(rec.data(), (size_t)L);
        auto range = seen.equal_range(h);
        for (auto it = range.first; it != range.second; ++it)
            if (std::memcmp(out.data() + (size_t)it->second * L, rec.data(), (size_t)L) == 0) return;
        seen.emplace(h, n);
        out.insert(out.end(), rec.begin(), rec.end());
        ++n;
    };

    std::string line;
    char buf[1 << 16];
    bool eof = false;
    while (!eof && !bad) {
        line.clear();
        // read one full line of arbitrary length
        for (;;) {
            if (!std::fgets(buf, sizeof buf, f)) { eof = true; break; }
            line.append(buf);
            if (!line.empty() && line.back() == '\n') break;
        }
        if (eof && line.empty()) break;
        ++line_no;
        if (!line.empty() && line[0] == '>') { flush(); fill = 0; continue; }
        bool any = false;
        for (char c : line) {
            if (c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
            any = true;
            if (fill < 0) fill = 0;             // headerless line: its own record
            if (fill < L) rec[(size_t)fill] = lut[(unsigned char)c];
            ++fill;
        }
        (void)any;
    }
    if (!bad) flush();
    std::fclose(f);
    if (bad) return -2;
    if (n == 0) { set_error("%s: no sequences found", path); return -3; }
    return n;
}

}  // namespace dca
