/* check_exp_glibc.c - host twin of reseek_b200/csrc/exp_glibc.cuh: the operation sequence of glibc's __exp_fma restated with
 * explicit fma() calls, compared bit for bit with this host's libm exp() on (a) every argument the DSS density sums can form
 * from float distances on a fine grid and (b) random doubles in [-1100, 1].  Build & run: tools/check_exp_glibc.sh
 * Exit status 0 = no mismatch. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RSK_HOSTDEV
#define RSK_FMA(a, b, c) fma((a), (b), (c))
#include "../reseek_b200/csrc/exp_glibc.cuh"

static uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rng(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(int argc, char **argv)
{
	const uint64_t nrand = argc > 1 ? strtoull(argv[1], 0, 10) : 200000000ull;
	uint64_t bad = 0, n = 0;
	/* (a) -Dist/20 with Dist = sqrtf(d2), d2 a float: every float d2 in [0, 4e4] on a stride of its bit pattern */
	for (uint32_t u = 0; u <= 0x471c4000u /* 40000.0f */; u += 7) {
		float d2; memcpy(&d2, &u, 4);
		const double Dist = sqrtf(d2);
		const double x = -Dist / 20.0;
		const double a = exp(x), b = rsk_exp_glibc(x);
		++n;
		if (bits(a) != bits(b) && bad++ < 10)
			printf("mismatch x=%a libm=%a ours=%a\n", x, a, b);
	}
	/* (b) random doubles */
	for (uint64_t k = 0; k < nrand; ++k) {
		const double u = (double)(rng() >> 11) / 9007199254740992.0;
		const double x = (k & 7) == 0 ? -1100.0 * u + 1.0 : -60.0 * u;
		const double a = exp(x), b = rsk_exp_glibc(x);
		++n;
		if (bits(a) != bits(b) && bad++ < 10)
			printf("mismatch x=%a libm=%a ours=%a\n", x, a, b);
	}
	/* (c) edge arguments */
	const double edge[] = {0.0, -0.0, -1e-300, -0x1p-54, -0x1p-55, -511.999, -512.0, -512.5, -700.0, -708.0, -708.5, -740.0, -745.2, -746.0, -800.0, -1e308};
	for (unsigned k = 0; k < sizeof(edge) / sizeof(edge[0]); ++k) {
		const double a = exp(edge[k]), b = rsk_exp_glibc(edge[k]);
		++n;
		if (bits(a) != bits(b) && bad++ < 10)
			printf("mismatch x=%a libm=%a ours=%a\n", edge[k], a, b);
	}
	printf("%llu arguments, %llu mismatches\n", (unsigned long long)n, (unsigned long long)bad);
	return bad ? 1 : 0;
}
