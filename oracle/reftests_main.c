/*
 * reftests_main.c — acceptance harness (test infrastructure, not shipped): runs the reference's OWN test
 * functions — compiled by oracle/Makefile from the sources where they lie under /root/reference/test
 * (bash_test.c:296, belt_test.c:822, bign_test.c:551, bign128_test.c:174, bign192_test.c, bign256_test.c,
 * math/ec_test.c:668, math/ecp_test.c:870) — in a process where libbee2_b200.so stands IN FRONT of the
 * unmodified reference library (link order -lbee2_b200 -lbee2ref_64). Every symbol libbee2_b200 exports
 * (bashF/bashHash*, bashPrg*, belt block/ECB/CTR/hash/DWP/CHE, bign sign/verify/keys/DH, ecMulA/ecAddMulA,
 * the bign128/192/256 forms) is taken from it — also by the reference's own internals (bignKeyWrap ->
 * ecMulA, beltCBC -> beltBlockEncr, ...) — and everything else from the reference. What a bee2
 * maintainer would run to accept the drop-in.
 *
 *   reftests_b200 [names...]     names: bash belt bign bign128 bign192 bign256 ec ecp (default: all)
 *   B2G_CPU_BELOW=<bytes>        route small one-shot calls to the stock library (overlay mode)
 * Prints "<name>Test: OK|Err" per test and b2g launch / forward counts; exit code = number of failures.
 */
#include <stdio.h>
#include <string.h>

typedef int bool_t;
extern bool_t bashTest(void), beltTest(void), bignTest(void), bign128Test(void), bign192Test(void),
	bign256Test(void), ecTest(void), ecpTest(void);
extern unsigned long long b2g_launch_count(void), b2g_forward_count(void);
extern int b2g_has_stock(void);
extern unsigned b2g_init(int);
extern const char* b2g_last_error(void);

static const struct { const char* name; bool_t (*fn)(void); } tests[] = {
	{"bash", bashTest}, {"belt", beltTest}, {"bign", bignTest}, {"bign128", bign128Test},
	{"bign192", bign192Test}, {"bign256", bign256Test}, {"ec", ecTest}, {"ecp", ecpTest}};

int main(int argc, char** argv)
{
	int fails = 0, ran = 0;
	size_t i;
	unsigned code = b2g_init(-1);
	printf("b2g_init: %u %s; stock libbee2 behind: %s\n", code, code ? b2g_last_error() : "ok",
		b2g_has_stock() ? "yes" : "no");
	for (i = 0; i < sizeof tests / sizeof tests[0]; ++i)
	{
		int want = argc < 2, a;
		for (a = 1; a < argc; ++a)
			want |= !strcmp(argv[a], tests[i].name);
		if (!want)
			continue;
		{
			const unsigned long long l0 = b2g_launch_count(), f0 = b2g_forward_count();
			const bool_t ok = tests[i].fn();
			printf("%sTest: %s  (gpu launches %llu, forwarded to stock %llu)\n", tests[i].name, ok ? "OK" : "Err",
				b2g_launch_count() - l0, b2g_forward_count() - f0);
			fflush(stdout);
			fails += !ok, ++ran;
		}
	}
	printf("ran %d, failed %d\n", ran, fails);
	return fails;
}
