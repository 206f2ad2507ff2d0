#!/usr/bin/env python3
"""Generates bee2_b200/csrc/gfp_asm.cuh: the carry-chain primitives of the GF(p) code.

Every primitive is ONE PTX asm statement (so nothing can clobber CC.CF inside a chain) with a
portable C twin under ``#ifndef __CUDA_ARCH__`` — the twin lets tests/host_gfp_test.cu run the
field / point logic built on top of these primitives on the CPU, where there is no GPU.

    python tools/gen_gfp_asm.py > bee2_b200/csrc/gfp_asm.cuh
"""

LIMBS = (8, 12, 16)          # bign levels 128 / 192 / 256: p has 32 * N bits


def ops(names):
    return ", ".join(names)


def emit_add_sub(n, sub):
    """r = a +/- b over n limbs; returns the carry (0/1) or the borrow MASK (0 / 0xFFFFFFFF)"""
    name = "sub_n" if sub else "add_n"
    op0, opc, opl = ("sub.cc.u32", "subc.cc.u32", "subc.u32") if sub else ("add.cc.u32", "addc.cc.u32", "addc.u32")
    lines = []
    for i in range(n):
        lines.append(f'"{op0 if i == 0 else opc} %{i}, %{n + 1 + i}, %{2 * n + 1 + i};\\n\\t"')
    lines.append(f'"{opl} %{n}, 0, 0;"')
    outs = ops([f'"=&r"(r[{i}])' for i in range(n)] + ['"=&r"(c)'])
    ins = ops([f'"r"(a[{i}])' for i in range(n)] + [f'"r"(b[{i}])' for i in range(n)])
    body = "\n\t\t".join(lines)
    if sub:
        host = (f"\tu64 w = 0;\n\tfor (int i = 0; i < {n}; ++i)\n\t{{\n\t\tconst u64 d = (u64)a[i] - b[i] - w;\n"
                f"\t\tr[i] = (u32)d, w = (d >> 32) & 1;\n\t}}\n\tc = (u32)0 - (u32)w;")
    else:
        host = (f"\tu64 w = 0;\n\tfor (int i = 0; i < {n}; ++i)\n\t{{\n\t\tw += (u64)a[i] + b[i];\n"
                f"\t\tr[i] = (u32)w, w >>= 32;\n\t}}\n\tc = (u32)w;")
    return f"""template <> GFP_HD u32 {name}<{n}>(u32* r, const u32* a, const u32* b)
{{
	u32 c;
#ifdef __CUDA_ARCH__
	asm({body}
		: {outs}
		: {ins});
#else
{host}
#endif
	return c;
}}
"""


def emit_add_cin(n):
    """r = a + b + cin over n limbs (cin = 0/1), the carry out is dropped (caller proves it 0)"""
    lines = [f'"{{\\n\\t.reg .u32 t;\\n\\tadd.cc.u32 t, %{3 * n}, 0xFFFFFFFF;\\n\\t"']
    for i in range(n):
        lines.append(f'"{"addc.cc.u32" if i < n - 1 else "addc.u32"} %{i}, %{n + i}, %{2 * n + i};\\n\\t{"}" if i == n - 1 else ""}"')
    outs = ops([f'"=&r"(r[{i}])' for i in range(n)])
    ins = ops([f'"r"(a[{i}])' for i in range(n)] + [f'"r"(b[{i}])' for i in range(n)] + ['"r"(cin)'])
    body = "\n\t\t".join(lines)
    return f"""template <> GFP_HD void add_n_cin<{n}>(u32* r, const u32* a, const u32* b, u32 cin)
{{
#ifdef __CUDA_ARCH__
	asm({body}
		: {outs}
		: {ins});
#else
	u64 w = cin;
	for (int i = 0; i < {n}; ++i)
	{{
		w += (u64)a[i] + b[i];
		r[i] = (u32)w, w >>= 32;
	}}
#endif
}}
"""


def emit_inc_dec(n, dec):
    """r +/-= x (one limb) over n limbs in place; returns carry (0/1) / borrow mask"""
    name = "dec_n" if dec else "inc_n"
    op0, opc, opl = ("sub.cc.u32", "subc.cc.u32", "subc.u32") if dec else ("add.cc.u32", "addc.cc.u32", "addc.u32")
    lines = [f'"{op0} %0, %0, %{n + 1};\\n\\t"']
    for i in range(1, n):
        lines.append(f'"{opc} %{i}, %{i}, 0;\\n\\t"')
    lines.append(f'"{opl} %{n}, 0, 0;"')
    outs = ops([f'"+r"(r[{i}])' for i in range(n)] + ['"=&r"(c)'])
    body = "\n\t\t".join(lines)
    if dec:
        host = (f"\tu64 w = x;\n\tfor (int i = 0; i < {n}; ++i)\n\t{{\n\t\tconst u64 d = (u64)r[i] - w;\n"
                f"\t\tr[i] = (u32)d, w = (d >> 32) & 1;\n\t}}\n\tc = (u32)0 - (u32)w;")
    else:
        host = (f"\tu64 w = x;\n\tfor (int i = 0; i < {n}; ++i)\n\t{{\n\t\tw += r[i];\n"
                f"\t\tr[i] = (u32)w, w >>= 32;\n\t}}\n\tc = (u32)w;")
    return f"""template <> GFP_HD u32 {name}<{n}>(u32* r, u32 x)
{{
	u32 c;
#ifdef __CUDA_ARCH__
	asm({body}
		: {outs}
		: "r"(x));
#else
{host}
#endif
	return c;
}}
"""


def emit_mad_row(l, top):
    """acc[0..2l) += sum_k a[2k] * b << (64 k); the carry out goes into acc[2l] (or is dropped: _top)"""
    name = "mad_row_top" if top else "mad_row"
    nacc = 2 * l + (0 if top else 1)
    lines = []
    for k in range(l):
        a_idx = nacc + k
        b_idx = nacc + l
        lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
        last = top and k == l - 1
        hi = "madc.hi.u32" if last else "madc.hi.cc.u32"
        lines.append(f'"{lo} %{2 * k}, %{a_idx}, %{b_idx}, %{2 * k};\\n\\t"')
        lines.append(f'"{hi} %{2 * k + 1}, %{a_idx}, %{b_idx}, %{2 * k + 1};{"" if last else "\\n\\t"}"')
    if not top:
        lines.append(f'"addc.u32 %{2 * l}, %{2 * l}, 0;"')
    outs = ops([f'"+r"(acc[{i}])' for i in range(nacc)])
    ins = ops([f'"r"(a[{2 * k}])' for k in range(l)] + ['"r"(b)'])
    body = "\n\t\t".join(lines)
    carry = "" if top else f"\tacc[{2 * l}] += (u32)w;\n"
    return f"""template <> GFP_HD void {name}<{l}>(u32* acc, const u32* a, u32 b)
{{
#ifdef __CUDA_ARCH__
	asm({body}
		: {outs}
		: {ins});
#else
	u64 w = 0;
	for (int k = 0; k < {l}; ++k)
	{{
		const u64 p = (u64)a[2 * k] * b;
		w += (u64)acc[2 * k] + (u32)p;
		acc[2 * k] = (u32)w, w >>= 32;
		w += (u64)acc[2 * k + 1] + (p >> 32);
		acc[2 * k + 1] = (u32)w, w >>= 32;
	}}
{carry}#endif
}}
"""


def emit_mad_diag(n):
    """t[0..2n) += sum_i a[i]^2 << (64 i); the final carry is dropped (caller proves it 0)"""
    lines = []
    for i in range(n):
        lo = "mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32"
        hi = "madc.hi.u32" if i == n - 1 else "madc.hi.cc.u32"
        lines.append(f'"{lo} %{2 * i}, %{2 * n + i}, %{2 * n + i}, %{2 * i};\\n\\t"')
        lines.append(f'"{hi} %{2 * i + 1}, %{2 * n + i}, %{2 * n + i}, %{2 * i + 1};{"" if i == n - 1 else "\\n\\t"}"')
    outs = ops([f'"+r"(t[{i}])' for i in range(2 * n)])
    ins = ops([f'"r"(a[{i}])' for i in range(n)])
    body = "\n\t\t".join(lines)
    return f"""template <> GFP_HD void mad_diag<{n}>(u32* t, const u32* a)
{{
#ifdef __CUDA_ARCH__
	asm({body}
		: {outs}
		: {ins});
#else
	u64 w = 0;
	for (int i = 0; i < {n}; ++i)
	{{
		const u64 p = (u64)a[i] * a[i];
		w += (u64)t[2 * i] + (u32)p;
		t[2 * i] = (u32)w, w >>= 32;
		w += (u64)t[2 * i + 1] + (p >> 32);
		t[2 * i + 1] = (u32)w, w >>= 32;
	}}
#endif
}}
"""


def main():
    out = ["""// gfp_asm.cuh — GENERATED by tools/gen_gfp_asm.py, do not edit.
//
// Carry-chain primitives for the GF(p) arithmetic of gfp.cuh (p = 2^(32 N) - c, N = 8 / 12 / 16
// limbs for bign levels 128 / 192 / 256). Each chain is ONE asm statement so that nothing can
// clobber CC.CF in between; pure outputs are early-clobber ("=&r") because the statements write
// them before the last input is read. mad.lo.cc / madc.hi.cc pairs are fused by ptxas into one
// IMAD.WIDE.U32(.X). The `#else` branches are the portable C twins used only when a translation
// unit is compiled for the host (tests/host_gfp_test.cu); device code never takes them.
#pragma once
#include "common.cuh"

#define GFP_HD __host__ __device__ __forceinline__

// r = a + b, returns the carry (0/1)
template <int N> GFP_HD u32 add_n(u32* r, const u32* a, const u32* b);
// r = a - b, returns the borrow mask (0 / 0xFFFFFFFF)
template <int N> GFP_HD u32 sub_n(u32* r, const u32* a, const u32* b);
// r = a + b + cin (cin = 0/1), carry out dropped
template <int N> GFP_HD void add_n_cin(u32* r, const u32* a, const u32* b, u32 cin);
// r += x (one limb), returns the carry; r -= x, returns the borrow mask
template <int N> GFP_HD u32 inc_n(u32* r, u32 x);
template <int N> GFP_HD u32 dec_n(u32* r, u32 x);
// acc[0..2L) += sum_k a[2k] * b << (64 k) (a is read with stride 2); carry into acc[2L] / dropped
template <int L> GFP_HD void mad_row(u32* acc, const u32* a, u32 b);
template <int L> GFP_HD void mad_row_top(u32* acc, const u32* a, u32 b);
// t[0..2N) += sum_i a[i]^2 << (64 i), final carry dropped
template <int N> GFP_HD void mad_diag(u32* t, const u32* a);
"""]
    sizes = sorted(set(LIMBS) | {n // 2 for n in LIMBS} | {n // 2 + 1 for n in LIMBS})
    for n in sizes:
        out.append(emit_add_sub(n, False))
        out.append(emit_add_sub(n, True))
    for n in LIMBS:
        out.append(emit_add_cin(n - 1))
    for n in sorted(set(LIMBS) | {n - 1 for n in LIMBS}):
        out.append(emit_inc_dec(n, False))
        out.append(emit_inc_dec(n, True))
    for l in range(1, max(LIMBS) // 2 + 1):
        out.append(emit_mad_row(l, False))
        out.append(emit_mad_row(l, True))
    for n in LIMBS:
        out.append(emit_mad_diag(n))
    print("\n".join(out), end="")


if __name__ == "__main__":
    main()
