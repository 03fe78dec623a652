#!/usr/bin/env python3
"""Regenerate profiles/sass/ from the CURRENT build (no GPU needed): one SASS listing per representative kernel and
SUMMARY.txt with the Blackwell-specific / system-scope mnemonics of every native kernel.

    python bench/dump_sass.py            # after `python -m tutel_b200._build`
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'sass')
INTERESTING = re.compile(r'^(UTC|LDTM|STTM|UTMA|UBLK|UCGABAR|MEMBAR|FENCE|REDG|ATOMG|RED\.|SYNCS|.*\.SYS|.*STRONG|LDG\.E\.NA|STG\.E\.NA|HMMA|MULTIMEM)')

# listing file -> regular expression on the demangled kernel name (first match is written out)
LISTINGS = {
    'gemm_sm100_2cta_kmajor_bn256_bf16.sass.txt': r'gemm_sm100_kernel<2, false, false, 256, 2, false>',
    'gemm_sm100_2cta_mnmajor_bn256_bf16.sass.txt': r'gemm_sm100_kernel<2, true, true, 256, 2, false>',
    'gemm_sm100_1cta_kmajor_bn256_bf16.sass.txt': r'gemm_sm100_kernel<1, false, false, 256, 2, false>',
    'gemm_sm100_2cta_kmajor_bn256_fp8.sass.txt': r'gemm_sm100_kernel<2, false, false, 256, 1, false>',
    'gemm_sm100_2cta_kmajor_bn256_bf16_xact.sass.txt': r'gemm_sm100_kernel<2, false, false, 256, 2, true>',
    'encode_rows_bf16_push.sass.txt': r'encode_rows_kernel<__nv_bfloat16, true, 128>',
    'decode_rows_bf16.sass.txt': r'decode_rows_kernel<__nv_bfloat16, true>',
    'gate_grad_bf16.sass.txt': r'gate_grad_kernel<__nv_bfloat16, true>',
    'gate_route_bf16_vpt1.sass.txt': r'gate_route_kernel<__nv_bfloat16, 1>',
    'route_finish_bf16.sass.txt': r'route_finish_kernel<__nv_bfloat16>',
    'gate_route_bwd_bf16_vpt1.sass.txt': r'gate_route_bwd_kernel<__nv_bfloat16, 1>',
    'colsum_bf16.sass.txt': r'colsum_kernel<__nv_bfloat16>',
    'p2p_push.sass.txt': r'p2p_push_kernel',
    'p2p_allreduce_oneshot_bf16.sass.txt': r'p2p_allreduce_oneshot_kernel<__nv_bfloat16>',
    'p2p_barrier.sass.txt': r'p2p_barrier_kernel',
    'p2p_stride_copy.sass.txt': r'p2p_stride_copy_kernel',
    'quantize_rows_bf16.sass.txt': r'quantize_rows_kernel<__nv_bfloat16>',
    'skinny_ffn_f32.sass.txt': r'skinny_ffn_kernel<float>',
    'gemm_mx_2cta_bn256.sass.txt': r'mx_gemm_kernel<2, 256>',
    'mx_quantize_bf16.sass.txt': r'mx_quantize_kernel<__nv_bfloat16>',
}


def main():
    so = sorted(glob.glob(os.path.join(ROOT, 'tutel_b200', '_C*.so')))
    if not so:
        sys.exit('build the extension first: python -m tutel_b200._build')
    sass = subprocess.run(['cuobjdump', '-sass', so[0]], capture_output=True, text=True, check=True).stdout
    os.makedirs(OUT, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, '*.sass.txt')):
        os.remove(f)
    kernels = []          # (demangled name, body)
    cur_name, cur = None, []
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if cur_name:
                kernels.append((cur_name, cur))
            cur_name, cur = m.group(1), []
        elif cur_name:
            cur.append(line)
    if cur_name:
        kernels.append((cur_name, cur))
    names = subprocess.run(['c++filt'], input='\n'.join(k for k, _ in kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    written = set()
    for (mangled, body), name in zip(kernels, names):
        short = name.replace('tb::(anonymous namespace)::', '').replace('(anonymous namespace)::', '').replace('void ', '')
        short = re.sub(r'\((?:[^()]|\([^()]*\))*\)\s*$', '', short)          # drop the argument list
        ops = collections.Counter()
        n = 0
        for line in body:
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)', line)
            if not m:
                continue
            n += 1
            if INTERESTING.match(m.group(1)):
                ops[m.group(1)] += 1
        rows.append((short, n, ops))
        for fname, pat in LISTINGS.items():
            if fname not in written and re.search(re.escape(pat), short):
                with open(os.path.join(OUT, fname), 'w') as f:
                    f.write('// %s\n// (cuobjdump -sass %s)\n' % (name, os.path.basename(so[0])))
                    f.write('\n'.join(body) + '\n')
                written.add(fname)
    with open(os.path.join(OUT, 'SUMMARY.txt'), 'w') as f:
        f.write('# Blackwell-specific / system-scope SASS mnemonics per kernel (bench/dump_sass.py on the current build, sm_100a)\n'
                '# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,\n'
                '# *.STRONG.SYS / MEMBAR.ALL.SYS / REDG...SYS = system-scope acquire / release on peer-mapped flags and counters\n')
        for short, n, ops in sorted(rows):
            f.write('%-96s %6d instr  %s\n' % (short[:96], n, ' '.join('%s=%d' % kv for kv in sorted(ops.items()))))
    missing = sorted(set(LISTINGS) - written)
    print('wrote %d listings + SUMMARY.txt (%d kernels)%s' % (len(written), len(rows), '; no match for: %s' % missing if missing else ''))


if __name__ == '__main__':
    main()
