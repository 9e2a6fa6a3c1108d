// gfa.cuh -- GFA1 + FASTA persistence of the graph, straight from the exported arrays (host code).
//
// Writer: the format of write_to_gfa1 (reference src/Graphs.jl:892-926) byte for byte: "H\tVN:Z:1.0", one
// "S\tseq<id>\t*\tLN:i:<len>\tUR:Z:<file>.fasta" line per node, the sequences in <file>.fasta (FASTX FASTA.Writer:
// ">seq<id>" + the sequence in lines of 70 bases), and one L line per stored link with source <= destination
// (sign -> orientation exactly as there; overlap = -dist, 0 for non-negative distances).  Nodes of length 0 are
// the reference's deleted nodes (empty_node, src/Graphs.jl:116-118, :459-466) and are skipped.
// Loader: load_from_gfa1! is unfinished upstream (it raises on the first GFA line, src/Graphs.jl:943-958); this
// one reads what the writer writes: node ids from the "seq<id>" names (gaps = deleted nodes), and per L line the
// inverse of the writer's mapping followed by add_link! (src/Graphs.jl:473-476: push (s, d) to |s|, (d, s) to |d|).
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

struct GfaGraphHost {
    std::vector<u64> node_off{0};
    std::vector<u8> bases;
    std::vector<u64> link_row{0};
    std::vector<i64> tri;
};

static int gfa1_write(const char *filename, u64 nn, const u64 *node_off, const u8 *bases, const u64 *link_row, const i64 *tri) {
    if (!filename || (nn && (!node_off || !link_row))) return GG_ERR_ARG;
    const std::string fa_name = std::string(filename) + ".fasta", gfa_name = std::string(filename) + ".gfa";
    FILE *gfa = fopen(gfa_name.c_str(), "w");
    if (!gfa) return GG_ERR_ARG;
    FILE *fa = fopen(fa_name.c_str(), "w");
    if (!fa) { fclose(gfa); return GG_ERR_ARG; }
    std::vector<char> big(1 << 20);
    setvbuf(gfa, nullptr, _IOFBF, 1 << 20);
    setvbuf(fa, big.data(), _IOFBF, big.size());
    fputs("H\tVN:Z:1.0\n", gfa);
    for (u64 i = 0; i < nn; ++i) {
        const u64 lo = node_off[i], hi = node_off[i + 1];
        if (hi == lo) continue;  // deleted node
        fprintf(gfa, "S\tseq%llu\t*\tLN:i:%llu\tUR:Z:%s\n", (unsigned long long)(i + 1), (unsigned long long)(hi - lo), fa_name.c_str());
        fprintf(fa, ">seq%llu\n", (unsigned long long)(i + 1));
        for (u64 p = lo; p < hi; p += 70) {
            fwrite(bases + p, 1, (size_t)((hi - p) < 70 ? (hi - p) : 70), fa);
            fputc('\n', fa);
        }
    }
    int rc = GG_OK;
    if (fclose(fa) != 0) rc = GG_ERR_ARG;
    for (u64 i = 0; i < nn; ++i)
        for (u64 q = link_row[i]; q < link_row[i + 1]; ++q) {
            const i64 s = tri[3 * q], d = tri[3 * q + 1], dist = tri[3 * q + 2];
            if (s > d) continue;
            fprintf(gfa, "L\tseq%lld\t%c\tseq%lld\t%c\t%lldM\n", (long long)(s > 0 ? s : -s), s > 0 ? '-' : '+', (long long)(d > 0 ? d : -d),
                    d > 0 ? '+' : '-', (long long)(dist < 0 ? -dist : 0));
        }
    if (fclose(gfa) != 0) rc = GG_ERR_ARG;
    return rc;
}

static bool gfa_seq_id(const char *s, u64 *id) {   // "seq<digits>"
    if (strncmp(s, "seq", 3) != 0 || s[3] < '0' || s[3] > '9') return false;
    u64 v = 0;
    const char *p = s + 3;
    for (; *p >= '0' && *p <= '9'; ++p) v = v * 10 + (u64)(*p - '0');
    *id = v;
    return v > 0 && (*p == 0 || *p == '\t' || *p == ' ' || *p == '\n' || *p == '\r');
}
static int gfa1_read(const char *gfa_file, const char *fasta_file, GfaGraphHost &g, std::string &err) {
    FILE *fa = fopen(fasta_file, "r");
    if (!fa) { err = std::string("cannot open ") + fasta_file; return GG_ERR_ARG; }
    std::vector<std::string> seqs;   // by node id - 1
    {
        std::vector<char> line(1 << 16);
        u64 cur = 0;
        while (fgets(line.data(), (int)line.size(), fa)) {
            size_t n = strlen(line.data());
            while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
            if (line[0] == '>') {
                if (!gfa_seq_id(line.data() + 1, &cur)) { fclose(fa); err = std::string("FASTA record name is not seq<id>: ") + line.data(); return GG_ERR_ARG; }
                if (seqs.size() < cur) seqs.resize(cur);
                if (!seqs[cur - 1].empty()) { fclose(fa); err = "duplicate FASTA record " + std::string(line.data()); return GG_ERR_ARG; }
            } else if (n) {
                if (cur == 0) { fclose(fa); err = "sequence data before the first FASTA header"; return GG_ERR_ARG; }
                seqs[cur - 1].append(line.data(), n);
            }
        }
        fclose(fa);
    }
    const u64 nn = seqs.size();
    std::vector<std::vector<i64>> links(nn);
    FILE *gfa = fopen(gfa_file, "r");
    if (!gfa) { err = std::string("cannot open ") + gfa_file; return GG_ERR_ARG; }
    {
        std::vector<char> line(1 << 16);
        while (fgets(line.data(), (int)line.size(), gfa)) {
            if (line[0] != 'L') continue;   // H and S lines carry nothing the FASTA file does not
            char a[64], b[64], oa, ob;
            long long ov = 0;
            if (sscanf(line.data(), "L\t%63s\t%c\t%63s\t%c\t%lldM", a, &oa, b, &ob, &ov) != 5) { fclose(gfa); err = std::string("malformed L line: ") + line.data(); return GG_ERR_ARG; }
            u64 ia = 0, ib = 0;
            if (!gfa_seq_id(a, &ia) || !gfa_seq_id(b, &ib) || ia > nn || ib > nn || (oa != '+' && oa != '-') || (ob != '+' && ob != '-')) {
                fclose(gfa); err = std::string("L line names an unknown node: ") + line.data(); return GG_ERR_ARG;
            }
            const i64 s = oa == '-' ? (i64)ia : -(i64)ia, d = ob == '+' ? (i64)ib : -(i64)ib, dist = -(i64)ov;
            links[ia - 1].insert(links[ia - 1].end(), {s, d, dist});
            links[ib - 1].insert(links[ib - 1].end(), {d, s, dist});
        }
        fclose(gfa);
    }
    g = GfaGraphHost();
    for (u64 i = 0; i < nn; ++i) {
        g.bases.insert(g.bases.end(), seqs[i].begin(), seqs[i].end());
        g.node_off.push_back(g.bases.size());
        g.tri.insert(g.tri.end(), links[i].begin(), links[i].end());
        g.link_row.push_back(g.tri.size() / 3);
    }
    return GG_OK;
}
