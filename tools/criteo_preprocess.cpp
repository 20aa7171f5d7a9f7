// criteo_preprocess -- label-encode the sparse columns of a Criteo TSV (label, I1..I13, C1..C26).
//
// Counterpart of the reference's test/criteo_preprocess.cpp (label-encodes Criteo-1TB with an
// open-addressing map; --repeat inflates the vocabulary). Own implementation: one pass builds a
// per-column dictionary (first-seen order) while streaming the encoded CSV to stdout/--out; a
// `meta` file with the per-column vocabulary sizes is written next to it (what the benchmark
// scripts read to size the tables).
//
//   g++ -O2 -std=c++17 tools/criteo_preprocess.cpp -o criteo_preprocess
//   ./criteo_preprocess --in day_0 --out train.csv --meta meta [--repeat 2]
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

int main(int argc, char** argv) {
    const char* in = nullptr; const char* out = nullptr; const char* meta = nullptr;
    int repeat = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--in")) in = argv[i + 1];
        else if (!strcmp(argv[i], "--out")) out = argv[i + 1];
        else if (!strcmp(argv[i], "--meta")) meta = argv[i + 1];
        else if (!strcmp(argv[i], "--repeat")) repeat = atoi(argv[i + 1]);
    }
    FILE* fi = in ? fopen(in, "r") : stdin;
    FILE* fo = out ? fopen(out, "w") : stdout;
    if (!fi || !fo) { fprintf(stderr, "cannot open input/output\n"); return 1; }
    const int ND = 13, NS = 26;
    std::vector<std::unordered_map<std::string, uint64_t>> dict(NS);
    fprintf(fo, "label");
    for (int i = 1; i <= ND; ++i) fprintf(fo, ",I%d", i);
    for (int i = 1; i <= NS; ++i) fprintf(fo, ",C%d", i);
    fprintf(fo, "\n");
    char* line = nullptr; size_t cap = 0; ssize_t len;
    uint64_t rows = 0;
    while ((len = getline(&line, &cap, fi)) > 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        std::vector<std::string> f;
        const char* p = line;
        while (true) {
            const char* t = strchr(p, '\t');
            if (!t) { f.emplace_back(p); break; }
            f.emplace_back(p, t - p);
            p = t + 1;
        }
        if ((int)f.size() < 1 + ND + NS) f.resize(1 + ND + NS);
        for (int r = 0; r < repeat; ++r) {
            fprintf(fo, "%s", f[0].empty() ? "0" : f[0].c_str());
            for (int i = 0; i < ND; ++i) fprintf(fo, ",%s", f[1 + i].empty() ? "0" : f[1 + i].c_str());
            for (int i = 0; i < NS; ++i) {
                std::string key = f[1 + ND + i];
                if (r) key += "#" + std::to_string(r);      // --repeat: distinct ids per replica inflate the vocabulary
                auto it = dict[i].find(key);
                uint64_t id = it == dict[i].end() ? (dict[i][key] = dict[i].size()) : it->second;
                fprintf(fo, ",%llu", (unsigned long long)id);
            }
            fprintf(fo, "\n");
            ++rows;
        }
    }
    free(line);
    if (meta) {
        FILE* fm = fopen(meta, "w");
        for (int i = 0; i < NS; ++i) fprintf(fm, "C%d %llu\n", i + 1, (unsigned long long)dict[i].size());
        fclose(fm);
    }
    fprintf(stderr, "criteo_preprocess: %llu rows\n", (unsigned long long)rows);
    if (in) fclose(fi);
    if (out) fclose(fo);
    return 0;
}
