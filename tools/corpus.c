/*
 * corpus.c -- deterministic synthetic corpora for tests and benchmarks
 * (TEST / BENCH INFRASTRUCTURE, not part of the product library).
 *
 * Implements the five generators of SURVEY.md section 8d.  Every generator is
 * segment addressable: jdc_fill(kind, first_byte_offset, dst, n) produces the
 * bytes [offset, offset+n) of an unbounded stream that is defined in
 * independent 4 MiB segments (PRNG re-seeded per segment from the corpus seed
 * and the segment index), so ranks / threads can generate their own slice.
 *
 *   kind 0 TEXT    Zipf(1.1) draws from 4096 pseudo words, punctuation, newlines
 *   kind 1 LOGS    timestamped service log lines from 64 templates
 *   kind 2 BINARY  16-byte little endian records (counter, u12, f32 walk, enum)
 *   kind 3 INCOMP  raw PRNG bytes
 *   kind 4 JSON    array of small objects with a TEXT message
 *   kind 5 MIXED   4 MiB segments cycling TEXT, BINARY, INCOMP (BASELINE config 2)
 *
 * PRNG: xorshift64* .
 */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#define API __attribute__((visibility("default")))
#define SEG ((uint64_t) 4 << 20)

static const uint64_t SEEDS[5] = {
	0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull, 0x94D049BB133111EBull,
	0x2545F4914F6CDD1Dull, 0xBF58476D1CE4E5B9ull
};

typedef struct { uint64_t s; } rng_t;

static inline uint64_t
rnd(rng_t* r)
{
	uint64_t x = r->s;
	x ^= x >> 12;
	x ^= x << 25;
	x ^= x >> 27;
	r->s = x;
	return x * 0x2545F4914F6CDD1Dull;
}

static inline uint32_t rnd_below(rng_t* r, uint32_t n) { return (uint32_t) ((rnd(r) >> 32) * (uint64_t) n >> 32); }

static void
seed_segment(rng_t* r, int kind, uint64_t seg)
{
	uint64_t s = SEEDS[kind] ^ (seg * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull);
	int i;
	if (s == 0) s = 1;
	r->s = s;
	for (i = 0; i < 4; i++) rnd(r);
}

/* ---- vocabulary ----------------------------------------------------------- */

#define NWORDS 4096
static char     vocab[NWORDS][13];
static uint8_t  vocab_len[NWORDS];
static uint32_t zipf_cdf[NWORDS];      /* 32-bit fixed point cumulative */
static int      vocab_ready;

static void
vocab_setup(void)
{
	/* rough English letter frequencies, per mille */
	static const char letters[] = "etaoinshrdlcumwfgypbvkjxqz";
	static const uint16_t freq[26] = { 127, 91, 82, 75, 70, 67, 63, 61, 60, 43, 40, 28, 28,
	                                   24, 24, 22, 20, 20, 19, 15, 10, 8, 2, 2, 1, 1 };
	uint32_t lcdf[26], tot = 0;
	double z[NWORDS], zs = 0, acc = 0;
	rng_t r;
	int i, k;

	if (vocab_ready) return;
	for (i = 0; i < 26; i++) { tot += freq[i]; lcdf[i] = tot; }
	r.s = SEEDS[0];
	for (i = 0; i < NWORDS; i++) {
		int len = 2 + (int) rnd_below(&r, 11);
		/* frequent words are short */
		if (i < 64 && len > 5) len = 2 + len % 4;
		for (k = 0; k < len; k++) {
			uint32_t v = rnd_below(&r, tot);
			int c = 0;
			while (lcdf[c] <= v) c++;
			vocab[i][k] = letters[c];
		}
		vocab[i][len] = 0;
		vocab_len[i] = (uint8_t) len;
	}
	for (i = 0; i < NWORDS; i++) { z[i] = 1.0 / pow((double) (i + 1), 1.1); zs += z[i]; }
	for (i = 0; i < NWORDS; i++) {
		acc += z[i] / zs;
		zipf_cdf[i] = acc >= 1.0 ? 0xffffffffu : (uint32_t) (acc * 4294967296.0);
	}
	zipf_cdf[NWORDS - 1] = 0xffffffffu;
	vocab_ready = 1;
}

static inline int
zipf_word(rng_t* r)
{
	uint32_t v = (uint32_t) (rnd(r) >> 32);
	int lo = 0, hi = NWORDS - 1;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (zipf_cdf[mid] < v) lo = mid + 1;
		else hi = mid;
	}
	return lo;
}

/* ---- per segment generators: fill buf[0..SEG) ------------------------------- */

typedef struct { uint8_t* p; uint8_t* end; } out_t;

static inline int
emit(out_t* o, const char* s, size_t n)
{
	size_t room = (size_t) (o->end - o->p);
	if (n > room) n = room;
	memcpy(o->p, s, n);
	o->p += n;
	return o->p < o->end;
}

static void
gen_text(uint8_t* buf, uint64_t seg)
{
	rng_t r;
	out_t o = { buf, buf + SEG };
	uint32_t nword = 0, punct = 5 + 0;
	seed_segment(&r, 0, seg);
	punct = 5 + rnd_below(&r, 9);
	for (;;) {
		int w = zipf_word(&r);
		if (!emit(&o, vocab[w], vocab_len[w])) break;
		nword++;
		if (nword % 14 == 0) {
			if (!emit(&o, "\n", 1)) break;
		} else if (--punct == 0) {
			punct = 5 + rnd_below(&r, 9);
			if (!emit(&o, (rnd(&r) & 1) ? ", " : ". ", 2)) break;
		} else {
			if (!emit(&o, " ", 1)) break;
		}
	}
}

static void
gen_logs(uint8_t* buf, uint64_t seg)
{
	static const char* levels[6] = { "INFO", "DEBUG", "WARN", "ERROR", "TRACE", "NOTICE" };
	static const char* verbs[16] = { "opened", "closed", "accepted", "rejected", "flushed", "retried",
		"scheduled", "evicted", "committed", "aborted", "loaded", "stored", "resolved", "queued",
		"dropped", "merged" };
	static const char* objs[8] = { "connection", "segment", "transaction", "cache entry", "request",
		"partition", "snapshot", "lease" };
	rng_t r;
	out_t o = { buf, buf + SEG };
	char line[256];
	/* every segment continues a monotonic clock: ~125 ms per line average */
	uint64_t ms = seg * 4000000ull;
	seed_segment(&r, 1, seg);
	for (;;) {
		uint32_t t, v, ob;
		int n;
		ms += rnd_below(&r, 251);
		t = (uint32_t) (ms % 86400000ull);
		v = rnd_below(&r, 16);
		ob = rnd_below(&r, 8);
		n = snprintf(line, sizeof(line),
			"2026-10-18T%02u:%02u:%02u.%03uZ host-%02u %s svc%u[%u]: %s %s id=%u addr=0x%08x took %u us\n",
			t / 3600000u, (t / 60000u) % 60u, (t / 1000u) % 60u, t % 1000u,
			rnd_below(&r, 32), levels[rnd_below(&r, 6)], rnd_below(&r, 8), 1000 + rnd_below(&r, 50),
			verbs[v], objs[ob], rnd_below(&r, 100000), (uint32_t) rnd(&r), rnd_below(&r, 50000));
		if (!emit(&o, line, (size_t) n)) break;
	}
}

static void
gen_binary(uint8_t* buf, uint64_t seg)
{
	rng_t r;
	uint64_t i, nrec = SEG / 16;
	uint32_t counter = (uint32_t) (seg * nrec);
	float walk = 100.0f;
	seed_segment(&r, 2, seg);
	for (i = 0; i < nrec; i++) {
		uint8_t* p = buf + i * 16;
		uint32_t u = rnd_below(&r, 4096);
		uint16_t en = (uint16_t) rnd_below(&r, 8), zero = 0;
		uint32_t c = counter++;
		walk += ((float) rnd_below(&r, 1001) - 500.0f) / 1000.0f;
		memcpy(p, &c, 4);
		memcpy(p + 4, &u, 4);
		memcpy(p + 8, &walk, 4);
		memcpy(p + 12, &en, 2);
		memcpy(p + 14, &zero, 2);
	}
}

static void
gen_incomp(uint8_t* buf, uint64_t seg)
{
	rng_t r;
	uint64_t i;
	seed_segment(&r, 3, seg);
	for (i = 0; i < SEG; i += 8) {
		uint64_t v = rnd(&r);
		memcpy(buf + i, &v, 8);
	}
}

/* one JSON object into line[]; returns its length */
static int
json_object(rng_t* r, char* line, size_t cap, uint32_t id)
{
	static const char* tags[16] = { "alpha", "beta", "gamma", "delta", "prod", "dev", "eu", "us",
		"batch", "online", "gold", "silver", "new", "legacy", "mobile", "web" };
	int n, k, words;
	n = snprintf(line, cap, "{\"id\":%u,\"name\":\"user%u\",\"tags\":[\"%s\",\"%s\",\"%s\"],\"score\":%.6f,\"ts\":%llu,\"msg\":\"",
		id, rnd_below(r, 1000000), tags[rnd_below(r, 16)], tags[rnd_below(r, 16)], tags[rnd_below(r, 16)],
		(double) rnd_below(r, 100000000) / 1000000.0,
		(unsigned long long) (1792300000000ull + (rnd(r) >> 40)));
	words = 5 + (int) rnd_below(r, 16);
	for (k = 0; k < words; k++) {
		int w = zipf_word(r);
		if (k) line[n++] = ' ';
		memcpy(line + n, vocab[w], vocab_len[w]);
		n += vocab_len[w];
	}
	memcpy(line + n, "\"}", 2);
	return n + 2;
}

static void
gen_json(uint8_t* buf, uint64_t seg)
{
	rng_t r;
	out_t o = { buf, buf + SEG };
	char line[512];
	uint32_t id = (uint32_t) (seg * 20000u);
	seed_segment(&r, 4, seg);
	emit(&o, "[", 1);
	for (;;) {
		int n = json_object(&r, line, sizeof(line), id++);
		if (!emit(&o, line, (size_t) n)) break;
		if (!emit(&o, ",\n", 2)) break;
	}
}

static void
gen_segment(int kind, uint64_t seg, uint8_t* buf)
{
	vocab_setup();
	if (kind == 5) {
		static const int cyc[3] = { 0, 2, 3 };
		kind = cyc[seg % 3];
	}
	switch (kind) {
		case 0: gen_text(buf, seg); break;
		case 1: gen_logs(buf, seg); break;
		case 2: gen_binary(buf, seg); break;
		case 3: gen_incomp(buf, seg); break;
		default: gen_json(buf, seg); break;
	}
}

/* bytes [offset, offset+n) of corpus `kind` */
API int
jdc_fill(int kind, uint64_t offset, uint8_t* dst, uint64_t n)
{
	uint8_t* tmp = NULL;
	if (kind < 0 || kind > 5) return -1;
	while (n) {
		uint64_t seg = offset / SEG, in = offset % SEG;
		uint64_t k = SEG - in < n ? SEG - in : n;
		if (in == 0 && k == SEG) {
			gen_segment(kind, seg, dst);
		} else {
			if (!tmp) tmp = malloc(SEG);
			if (!tmp) return -2;
			gen_segment(kind, seg, tmp);
			memcpy(dst, tmp + in, k);
		}
		dst += k;
		offset += k;
		n -= k;
	}
	free(tmp);
	return 0;
}

/*
 * One JSON record of exactly `size` bytes for BASELINE config 3: an array of
 * objects truncated to the record size and closed with ']'.
 */
API int
jdc_json_record(uint64_t index, uint8_t* dst, uint32_t size)
{
	rng_t r;
	char line[512];
	uint32_t pos = 0, id = (uint32_t) (index * 977u);
	vocab_setup();
	if (size < 2) return -1;
	r.s = SEEDS[4] ^ (index * 0xD6E8FEB86659FD93ull + 1);
	rnd(&r); rnd(&r);
	dst[pos++] = '[';
	while (pos < size - 1) {
		int n = json_object(&r, line, sizeof(line), id++);
		uint32_t room = size - 1 - pos;
		uint32_t k = (uint32_t) n < room ? (uint32_t) n : room;
		memcpy(dst + pos, line, k);
		pos += k;
		if (pos < size - 1) dst[pos++] = ',';
	}
	dst[size - 1] = ']';
	return 0;
}

/* record size for config 3: 4 KiB * 2^U[0,4) (log-uniform 4-64 KiB) */
API uint32_t
jdc_json_record_size(uint64_t index)
{
	rng_t r;
	double u;
	r.s = 0x8CB92BA72F3D8DD7ull ^ (index * 0x9E3779B97F4A7C15ull + 7);
	rnd(&r); rnd(&r);
	u = (double) (rnd(&r) >> 11) / 9007199254740992.0 * 4.0;
	return (uint32_t) (4096.0 * pow(2.0, u));
}
