/* rand_glibc.c -- srand()/rand() with glibc's default generator, without the lock.
 *
 * The calibration's emitters (emit_fast.c) spend most of their time inside rand(): glibc takes a lock
 * around every call.  These definitions produce the SAME sequence as glibc's srand()/rand() -- the
 * additive feedback generator of random_r.c, TYPE_3: 31 words, r[i] = r[i-31] + r[i-3], seeded with the
 * Lehmer sequence 16807 * x mod (2^31 - 1) and warmed up by 310 draws; rand() returns the sum >> 1 --
 * so `-seed` keeps selecting the same simulated reads as the CPU reference.  tests/test_emit_host.py
 * compares them with libc's for many seeds, and the gold tests (which fix -seed 42) depend on it.
 * The drop-in binary is single-threaded wherever rand() is used (calibrateQ.c, simulate code).
 */
#include <stdint.h>
#include <stdlib.h>

static int32_t g_r[34];
static int g_f = 3, g_b = 0;   /* front and rear positions, as after srand() */

static void seed_state(unsigned int seed)
{
	int i;
	int32_t word;
	if (seed == 0) seed = 1;
	g_r[0] = (int32_t)seed;
	word = (int32_t)seed;
	for (i = 1; i < 31; i++) {
		/* word = 16807 * word % 2147483647 without overflowing 31 bits (Schrage) */
		long int hi = word / 127773;
		long int lo = word % 127773;
		word = (int32_t)(16807 * lo - 2836 * hi);
		if (word < 0) word += 2147483647;
		g_r[i] = word;
	}
	g_f = 3; g_b = 0;
}

static inline int32_t step(void)
{
	uint32_t v = (uint32_t)g_r[g_f] + (uint32_t)g_r[g_b];
	g_r[g_f] = (int32_t)v;
	if (++g_f >= 31) g_f = 0;
	if (++g_b >= 31) g_b = 0;
	return (int32_t)(v >> 1);
}

static int g_seeded = 0;

void srand(unsigned int seed)
{
	int i;
	seed_state(seed);
	for (i = 0; i < 310; i++) (void)step();
	g_seeded = 1;
}

int rand(void)
{
	if (!g_seeded) srand(1);   /* like an unseeded libc generator */
	return (int)step();
}
