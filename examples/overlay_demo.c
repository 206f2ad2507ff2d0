/*
 * overlay_demo.c — a plain bee2 application (it only uses names and prototypes of include/bee2) that is
 * accelerated WITHOUT a source change, the packaging SURVEY.md §8b describes:
 *
 *   demo_stock    : linked with the stock libbee2 only                       (CPU, the reference's result)
 *   demo_linked   : linked  -lbee2_b200 -lbee2   (the engine in FRONT of stock libbee2)
 *   LD_PRELOAD=libbee2_b200.so demo_stock : the same binary as the first line, engine interposed at run time
 *
 * In the last two, every hot-path symbol (bashHash, beltCTR, beltECB*, beltHash, bign*, ecMulA ...) is
 * served by libbee2_b200.so; inputs it does not cover (here: bignVerify under a NON-standard parameter
 * block, and everything outside the hot path: beltCBCEncr, hexFrom ...) reach the stock library through
 * dlsym(RTLD_NEXT); with B2G_CPU_BELOW=<bytes> small one-shot calls do too. The three runs must print the
 * same digests (tests/test_gpu_reftests.py::test_overlay_demo).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include "../include/bee2_b200.h"   /* prototypes identical to include/bee2/crypto/{bash,belt,bign}.h */

/* outside the hot path: always the stock library's (belt.h:548-571, hex.h) */
extern err_t beltCBCEncr(void* dest, const void* src, size_t count, const octet key[], size_t len, const octet iv[16]);
extern err_t bignParamsVal(const bign_params* params);

static void hex(const char* what, const octet* p, size_t n)
{
	size_t i;
	printf("%-28s", what);
	for (i = 0; i < n; ++i)
		printf("%02X", p[i]);
	printf("\n");
}

int main(void)
{
	const octet* H = beltH();
	octet hash[64], mac[8], sig[48], priv[32], pub[64];
	octet* big = (octet*)malloc((size_t)8 << 20);
	bign_params params;
	err_t code;
	size_t i;
	u64 (*launches)(void) = (u64(*)(void))dlsym(RTLD_DEFAULT, "b2g_launch_count");
	u64 (*forwards)(void) = (u64(*)(void))dlsym(RTLD_DEFAULT, "b2g_forward_count");

	/* a 13-byte message (STB 34.101.77 A.3 style) and an 8 MiB one */
	if (bashHash(hash, 128, H, 13)) return 1;
	hex("bash256(13 B)", hash, 32);
	for (i = 0; i < ((size_t)8 << 20); ++i) big[i] = H[(i * 7 + (i >> 8)) & 255];
	if (bashHash(hash, 256, big, (size_t)8 << 20)) return 2;
	hex("bash512(8 MiB)", hash, 64);
	/* belt: CTR over the big buffer, hash of the ciphertext; CBC is not on the hot path -> stock */
	if (beltCTR(big, big, (size_t)8 << 20, H + 128, 32, H + 192)) return 3;
	if (beltHash(hash, big, (size_t)8 << 20)) return 4;
	hex("belt-hash(CTR(8 MiB))", hash, 32);
	if (beltCBCEncr(big, big, 4096, H + 128, 32, H + 192)) return 5;
	if (beltHash(hash, big, 4096)) return 6;
	hex("belt-hash(CBC(4 KiB))", hash, 32);
	if (beltDWPWrap(big, mac, big, 1 << 20, H, 32, H + 128, 32, H + 192)) return 7;
	hex("belt-DWP mac(1 MiB)", mac, 8);
	/* bign on the standard curve: keys, deterministic signature, verification */
	if (bignParamsStd(&params, "1.2.112.0.2.0.34.101.45.3.1")) return 8;
	memcpy(priv, H + 32, 32), priv[31] &= 0x7F;
	if (bignPubkeyCalc(pub, &params, priv)) return 9;
	hex("bign pubkey", pub, 64);
	if (beltHash(hash, H, 13)) return 10;
	{
		static const octet oid[] = {0x06, 0x09, 0x2A, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x1F, 0x51};
		if (bignSign2(sig, &params, oid, sizeof oid, hash, priv, 0, 0)) return 11;
		hex("bign sign2", sig, 48);
		code = bignVerify(&params, oid, sizeof oid, hash, sig, pub);
		printf("%-28s%u\n", "bign verify", code);
		sig[0] ^= 1;
		code = bignVerify(&params, oid, sizeof oid, hash, sig, pub);
		printf("%-28s%u\n", "bign verify (bad sig)", code);
		/* a parameter block that passes the structural checks but is not a standard curve: the engine
		   has no GPU path for it (ERR_NOT_IMPLEMENTED = 119 stand-alone); in front of stock libbee2 the
		   call is forwarded and stock answers as it always did */
		params.b[0] ^= 2;
		code = bignVerify(&params, oid, sizeof oid, hash, sig, pub);
		printf("%-28s%u\n", "bign verify (foreign curve)", code);
	}
	if (launches && forwards)
		fprintf(stderr, "engine present: %llu kernel launches, %llu calls forwarded to stock libbee2\n",
			(unsigned long long)launches(), (unsigned long long)forwards());
	else
		fprintf(stderr, "stock libbee2 only\n");
	free(big);
	return 0;
}
