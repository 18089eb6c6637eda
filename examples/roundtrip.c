/*
 * examples/roundtrip.c -- the reference's README loops (README.md:98-184 of Jpn666/jdeflate),
 * unchanged, linked against the B200 build: raw deflate -> inflate through deflator_* / inflator_*,
 * then gzip through the zstrm callbacks.  Nothing here knows about CUDA.
 *
 *   gcc -std=c99 examples/roundtrip.c -Iinclude -Ljdeflate_b200/lib -ljdeflate \
 *       -Wl,-rpath,$PWD/jdeflate_b200/lib -o roundtrip && ./roundtrip
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <jdeflate/deflator.h>
#include <jdeflate/inflator.h>
#include <jdeflate/zstrm.h>

struct sink { uint8* p; uintxx n, cap; };
struct feed { const uint8* p; uintxx n, pos; };

static intxx
put(const uint8* buffer, uintxx size, void* user)
{
	struct sink* s = user;
	if (s->n + size > s->cap) return -1;
	memcpy(s->p + s->n, buffer, size);
	s->n += size;
	return (intxx) size;
}

static intxx
get(uint8* buffer, uintxx size, void* user)
{
	struct feed* f = user;
	uintxx k = f->n - f->pos;
	if (k > size) k = size;
	memcpy(buffer, f->p + f->pos, k);
	f->pos += k;
	return (intxx) k;
}

int
main(void)
{
	const uintxx n = 3u << 20;
	uint8* src = malloc(n);
	uint8* cmp = malloc(n + n / 8 + 4096);
	uint8* out = malloc(n);
	uintxx i, clen = 0, olen = 0;
	uint32 seed = 12345;
	const struct JDEFLATEVersion version = jdeflate_getversion();

	printf("jdeflate %s\n", version.versionstring);
	if (!src || !cmp || !out) return 2;
	for (i = 0; i < n; i++) {                         /* compressible, not trivial */
		seed = seed * 1103515245u + 12345u;
		src[i] = (uint8) ("the quick brown fox jumps over the lazy dog "[(i + (seed >> 28)) % 44]);
	}

	/* ---- raw DEFLATE, small target windows (README.md:98-140) ---- */
	{
		TDeflator* d = deflator_create(0, 6, NULL);
		uint8 window[4096];
		eDEFLTResult r;
		if (d == NULL) { fprintf(stderr, "deflator_create failed (no CUDA device?)\n"); return 3; }
		deflator_setsrc(d, src, n);
		do {
			deflator_settgt(d, window, sizeof(window));
			r = deflator_deflate(d, DEFLT_END);
			memcpy(cmp + clen, window, deflator_tgtend(d));
			clen += deflator_tgtend(d);
		} while (r == DEFLT_TGTEXHSTD);
		if (r != DEFLT_OK) { fprintf(stderr, "deflate: %d error %u\n", (int) r, (unsigned) d->error); return 4; }
		deflator_destroy(d);
	}
	{
		TInflator* s = inflator_create(0, NULL);
		eINFLTResult r;
		if (s == NULL) return 3;
		inflator_setsrc(s, cmp, clen);
		do {
			inflator_settgt(s, out + olen, 65536 < n - olen ? 65536 : n - olen + 1);
			r = inflator_inflate(s, 1);
			olen += inflator_tgtend(s);
		} while (r == INFLT_TGTEXHSTD && olen < n);
		if (r != INFLT_OK && !(r == INFLT_TGTEXHSTD && olen == n)) {
			fprintf(stderr, "inflate: %d error %u\n", (int) r, (unsigned) s->error);
			return 5;
		}
		inflator_destroy(s);
	}
	if (olen != n || memcmp(src, out, n) != 0) { fprintf(stderr, "raw round trip differs\n"); return 6; }
	printf("raw:  %lu -> %lu bytes, round trip ok\n", (unsigned long) n, (unsigned long) clen);

	/* ---- gzip through the zstrm callbacks (README.md:186-331) ---- */
	{
		struct sink sk = { cmp, 0, n + n / 8 + 4096 };
		struct feed fd;
		const TZStrm* z = zstrm_create(ZSTRM_DEFLATE | ZSTRM_GZIP, 6, NULL);
		uint32 crc;
		if (z == NULL) return 3;
		zstrm_settargetfn(z, put, &sk);
		for (i = 0; i < n; i += 1u << 20) zstrm_deflate(z, src + i, (n - i) < (1u << 20) ? (n - i) : (1u << 20));
		zstrm_flush(z, 1);
		if (z->error) { fprintf(stderr, "zstrm deflate error %u\n", (unsigned) z->error); return 7; }
		crc = z->crc;
		zstrm_destroy(z);

		fd.p = cmp; fd.n = sk.n; fd.pos = 0;
		z = zstrm_create(ZSTRM_INFLATE, 0, NULL);
		if (z == NULL) return 3;
		zstrm_setsourcefn(z, get, &fd);
		olen = 0;
		memset(out, 0, n);
		while (z->state != ZSTRM_END) {
			uintxx got = zstrm_inflate(z, out + olen, (n - olen) < 100000 ? (n - olen) + 1 : 100000);
			olen += got;
			if (got == 0 && z->state != ZSTRM_END) break;
		}
		if (z->error || olen != n || memcmp(src, out, n) != 0 || z->crc != crc ||
		    crc != (zstrm_crc32update(0xffffffffu, src, n) ^ 0xffffffffu)) {
			fprintf(stderr, "gzip round trip: error %u, %lu bytes\n", (unsigned) z->error, (unsigned long) olen);
			return 8;
		}
		printf("gzip: %lu -> %lu bytes, crc %08x, round trip ok\n", (unsigned long) n, (unsigned long) sk.n, (unsigned) crc);
		zstrm_destroy(z);
	}
	free(src); free(cmp); free(out);
	return 0;
}
