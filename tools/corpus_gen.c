/* corpus_gen.c — deterministic synthetic corpora for tests and bench.py
 * (SURVEY.md §8d configs C1-C5).  Bench/test infrastructure, not product.
 * All randomness is splitmix64 with fixed seeds so the CPU oracle and the GPU
 * path always see identical bytes. */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

#define API __attribute__((visibility("default")))

typedef struct { uint64_t st; } Rng;
static inline uint64_t next64(Rng *r)
{
    r->st += 0x9E3779B97F4A7C15ull;
    uint64_t z = r->st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double nextf(Rng *r) { return (double)(next64(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t nextn(Rng *r, uint32_t n) { return (uint32_t)(((next64(r) >> 32) * (uint64_t)n) >> 32); }

/* C5: incompressible bytes, 8 per splitmix64 call, little endian */
API void corpus_random(uint64_t seed, uint8_t *out, size_t n)
{
    Rng r = { seed };
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t z = next64(&r); memcpy(out + i, &z, 8); }
    if (i < n) { uint64_t z = next64(&r); memcpy(out + i, &z, n - i); }
}

/* V8-style: one byte per call (low byte) */
API void corpus_random_bytewise(uint64_t seed, uint8_t *out, size_t n)
{
    Rng r = { seed };
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)(next64(&r) & 0xFF);
}

/* ---- English-like text (C1) ---- */
#define VOCAB 4096
typedef struct { char w[VOCAB][12]; uint8_t len[VOCAB]; } Vocab;

static void make_vocab(Vocab *v, uint64_t seed)
{
    /* cumulative English unigram weights (per mille-ish) for a..z */
    static const uint16_t wt[26] = { 82, 15, 28, 43, 127, 22, 20, 61, 70, 2, 8, 40, 24,
                                     67, 75, 19, 1, 60, 63, 91, 28, 10, 24, 2, 20, 1 };
    uint32_t cum[26], tot = 0;
    for (int i = 0; i < 26; i++) { tot += wt[i]; cum[i] = tot; }
    Rng r = { seed ^ 0x766F636162ull };
    for (int k = 0; k < VOCAB; k++) {
        /* frequent (low-rank) words are short */
        int len = 2 + (int)nextn(&r, k < 64 ? 3 : (k < 512 ? 6 : 9));
        v->len[k] = (uint8_t)len;
        for (int j = 0; j < len; j++) {
            uint32_t x = nextn(&r, tot);
            int c = 0;
            while (cum[c] <= x) c++;
            v->w[k][j] = (char)('a' + c);
        }
    }
}

static size_t gen_text(Rng *r, const Vocab *v, uint8_t *out, size_t n)
{
    const double lnV = log((double)VOCAB);
    size_t i = 0;
    int cap = 1;
    while (i < n) {
        uint32_t rank = (uint32_t)exp(nextf(r) * lnV);      /* floor(4096^u) - Zipf-like */
        if (rank >= VOCAB) rank = VOCAB - 1;
        int len = v->len[rank];
        for (int j = 0; j < len && i < n; j++) {
            char c = v->w[rank][j];
            if (cap && j == 0) c = (char)(c - 32);
            out[i++] = (uint8_t)c;
        }
        cap = 0;
        uint32_t s = nextn(r, 100);
        if (s < 85) { if (i < n) out[i++] = ' '; }
        else if (s < 90) { if (i < n) out[i++] = ','; if (i < n) out[i++] = ' '; }
        else if (s < 96) { if (i < n) out[i++] = '.'; if (i < n) out[i++] = ' '; cap = 1; }
        else { if (i < n) out[i++] = '\n'; }
    }
    return i;
}

API void corpus_text(uint64_t seed, uint8_t *out, size_t n)
{
    static Vocab v;
    make_vocab(&v, 0xB2000001ull);         /* one shared vocabulary, like one language */
    Rng r = { seed };
    gen_text(&r, &v, out, n);
}

/* ---- source-code-like: ~200 line templates with random identifiers/integers ---- */
static size_t gen_source(Rng *r, uint8_t *out, size_t n)
{
    static const char *tmpl[] = {
        "    let mut % = %.len() - $;\n", "    if % < % && % != $ {\n", "        return Err(%::new($));\n",
        "    }\n", "fn %(%: &mut %, %: usize) -> % {\n", "    for % in $..% {\n", "        %[%] = %[% + $] ^ %;\n",
        "    // TODO(%): handle % overflow when % > $\n", "    assert!(% <= %, \"% out of range: {}\", %);\n",
        "static const uint32_t %[$] = { $, $, $, $ };\n", "#include <%/%.h>\n", "}\n", "\n",
        "    %->% = (%_t *)malloc(sizeof(*%) * #);\n", "    while (% > # && %[% - #] == %) { %--; }\n",
        "    printf(\"%=%d %=%d\\n\", %, %);\n", "    pub fn %(&self) -> &% { &self.% }\n",
        "impl % for % {\n", "        let % = self.%.%(%, #)?;\n", "    match % { Some(%) => %, None => # }\n",
    };
    static const char *idents[] = {
        "buf", "len", "idx", "count", "state", "ctx", "node", "left", "right", "value", "key", "table",
        "offset", "cursor", "block", "stream", "writer", "reader", "config", "result", "error", "handle",
        "queue", "bucket", "rank", "suffix", "prefix", "symbol", "freq", "weight", "parent", "child",
    };
    const int NT = sizeof tmpl / sizeof *tmpl, NI = sizeof idents / sizeof *idents;
    size_t i = 0;
    while (i < n) {
        const char *t = tmpl[nextn(r, (uint32_t)NT)];
        for (; *t && i < n; t++) {
            if (*t == '%') {
                const char *id = idents[nextn(r, (uint32_t)NI)];
                for (; *id && i < n; id++) out[i++] = (uint8_t)*id;
                if (nextn(r, 4) == 0 && i < n) out[i++] = (uint8_t)('0' + nextn(r, 10));
            } else if (*t == '$') {
                char num[16];
                int k = snprintf(num, sizeof num, "%u", nextn(r, nextn(r, 3) == 0 ? 100000u : 64u));
                for (int j = 0; j < k && i < n; j++) out[i++] = (uint8_t)num[j];
            } else {
                out[i++] = (uint8_t)*t;
            }
        }
    }
    return i;
}

API void corpus_source(uint64_t seed, uint8_t *out, size_t n)
{
    Rng r = { seed };
    gen_source(&r, out, n);
}

/* ---- binary-like: LE u32 counters, small-alphabet tables, zero pages, random pages ---- */
static size_t gen_binary(Rng *r, uint8_t *out, size_t n)
{
    size_t i = 0;
    while (i < n) {
        uint32_t kind = nextn(r, 4);
        size_t len = 4096 * (1 + nextn(r, 16));
        if (len > n - i) len = n - i;
        if (kind == 0) {                 /* counters with a random stride */
            uint32_t v = (uint32_t)next64(r), stride = 1 + nextn(r, 8);
            for (size_t k = 0; k + 4 <= len; k += 4) { memcpy(out + i + k, &v, 4); v += stride; }
            for (size_t k = len & ~(size_t)3; k < len; k++) out[i + k] = 0;
        } else if (kind == 1) {          /* small-alphabet table */
            uint32_t sigma = 2 + nextn(r, 14);
            for (size_t k = 0; k < len; k++) out[i + k] = (uint8_t)(nextn(r, sigma) * 17);
        } else if (kind == 2) {          /* zero page */
            memset(out + i, 0, len);
        } else {                         /* random page */
            for (size_t k = 0; k < len; k++) out[i + k] = (uint8_t)next64(r);
        }
        i += len;
    }
    return i;
}

API void corpus_binary(uint64_t seed, uint8_t *out, size_t n)
{
    Rng r = { seed };
    gen_binary(&r, out, n);
}

/* C2: 4 MiB extents cycling text / source / binary */
API void corpus_mixed(uint64_t seed, uint8_t *out, size_t n)
{
    static Vocab v;
    make_vocab(&v, 0xB2000001ull);
    Rng r = { seed };
    const size_t EXT = (size_t)4 << 20;
    size_t i = 0;
    int k = 0;
    while (i < n) {
        size_t len = n - i < EXT ? n - i : EXT;
        if (k % 3 == 0) gen_text(&r, &v, out + i, len);
        else if (k % 3 == 1) gen_source(&r, out + i, len);
        else gen_binary(&r, out + i, len);
        i += len;
        k++;
    }
}
