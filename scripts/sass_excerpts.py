"""SASS evidence for libais_b200.so: mnemonic counts of the whole library and of the hot kernels, with excerpts
(profiles/r2_sass_excerpts.txt is this script's output). Needs cuobjdump; no GPU. Usage: sass_excerpts.py > out.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ais_benchmarks_b200", "libais_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        funcs[name].append(line.rstrip())


def mnemonic(line):
    body = re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", line)
    body = re.sub(r"^@!?U?P\d+\s+", "", body)
    return body.split()[0].rstrip(";").split(".")[0]


def counts(lines):
    return collections.Counter(mnemonic(l) for l in lines)


allc = collections.Counter()
for ls in funcs.values():
    allc.update(counts(ls))
keys = ["UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "MUFU", "HMMA",
        "UTMALDG"]
print("SASS evidence for libais_b200.so (cuobjdump -sass ais_benchmarks_b200/libais_b200.so; sm_100a), round 2 final build;")
print("regenerate with scripts/sass_excerpts.py.")
print("Whole library, selected mnemonics: " + ", ".join("%s x%d" % (k, allc.get(k, 0)) for k in keys))
print("(UTCHMMA = tcgen05.mma kind::tf32, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc,")
print(" UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, FFMA2/FADD2/FMUL2 = packed FP32, FMNMX3 = 3-input min/max;")
print(" no HMMA (legacy mma.sync) and no UTMALDG (tensor-map TMA: the bulk copies are 1-D) anywhere.)")


def show(title, pattern, anchor, before, after, note):
    for fn, ls in funcs.items():
        if re.search(pattern, fn):
            c = counts(ls)
            print("\n== %s\n   %s" % (title, fn))
            print("   " + ", ".join("%s x%d" % kv for kv in c.most_common(14)))
            idx = [i for i, l in enumerate(ls) if re.search(anchor, l)]
            if idx:
                i = idx[0]
                print("   excerpt: " + note)
                for l in ls[max(0, i - before):i + after]:
                    print("        " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
            return
    print("\n== %s: NOT FOUND (%s)" % (title, pattern))


show("logpdf_fast_kernel<8,false> (headline KDE kernel; two groups of 8 components per trip of the inner loop)",
     r"18logpdf_fast_kernelILi8ELb0E", r"VOTE\.ALL", 34, 14,
     "end of one group's screen (packed FFMA2 chains with scalar-broadcast operands, FMNMX3 minima), the vote, the branch:")
show("tc_logq_kernel<32> (tcgen05 proposal pass)", r"14tc_logq_kernelILi32E", r"UTCHMMA", 6, 10,
     "MMA issue (tcgen05.mma kind::tf32 from shared-memory descriptors into TMEM) and commit:")
show("tc_logq_kernel<32>, epilogue", r"14tc_logq_kernelILi32E", r"LDTM", 2, 26,
     "TMEM load (tcgen05.ld 32x32b.x32) and the FMNMX3 chunk maximum:")
show("logpdf_kernel<DIAG,32> (target pass of the DM step)", r"13logpdf_kernelILi1ELi32ELb1ELb0E", r"UBLKCP", 3, 8,
     "1-D TMA bulk copy of a staged chunk (cp.async.bulk + mbarrier expect-tx):")
show("logpdf_subset_kernel<8> (exact re-evaluation, self-planning launch)", r"20logpdf_subset_kernelILi8E", r"UBLKCP", 2, 6,
     "bulk copy in the strided work loop:")
show("gather_rows_vec_kernel", r"gather_rows_vec_kernel", r"LDG", 4, 6, "one 16-byte piece of a row per thread:")
