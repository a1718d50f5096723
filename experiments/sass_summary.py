"""SASS evidence per kernel of libukbb_fcn.so: tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), TMA (UTMALDG / UTMASTG), mbarrier
(SYNCS) and packed-FP32 mnemonics, from `cuobjdump -sass`.  usage: sass_summary.py all.sass [out_dir]
Writes a table to stdout and, with out_dir, the full listing of the kernels on the default BF16 path (gzip)."""
import collections, gzip, re, sys
path = sys.argv[1]
out_dir = sys.argv[2] if len(sys.argv) > 2 else None
KEYS = ["UTCHMMA", "UTCQMMA", "2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "FFMA2", "FADD2", "F2FP", "LDG", "STG", "LDS", "STS", "REDG"]
# kernels on the DEFAULT path (mode fp16x2: F16 = SPLIT = F8 = true template flags; PAIR = true for conv_halo and conv_group 64 -> 64)
DEFAULT = ["conv_first_tc_kernelILb1ELb1ELb1", "conv_group_kernelILi16ELi32ELi2ELb1ELb1ELb1ELb0", "conv_group_kernelILi32ELi32ELi1ELb1ELb1ELb1ELb0",
           "conv_group_kernelILi32ELi64ELi2ELb1ELb1ELb1ELb0", "conv_group_kernelILi64ELi64ELi1ELb1ELb1ELb1ELb1", "conv_tc_kernelILi64ELi128ELb1ELb1ELb1",
           "conv_tc_kernelILi64ELi256ELb1ELb1ELb1", "conv_halo_kernelILi32ELi128ELb0ELi0ELb1ELb1ELb1ELb1ELb1", "conv_halo_kernelILi32ELi256ELb0ELi0ELb1ELb1ELb1ELb1ELb1",
           "side_tc_kernelILb1ELb1ELb1", "head_ts_kernelILi4ELb1ELb1ELb1", "int_hist", "int_scan", "lut_kernel", "rescale_lut", "sel_hist0", "sel_histn",
           "sel_scan", "sel_final", "sel_init", "rescale_pad", "cc_stats_kernel", "convT_fp32", "lstm_point", "ao_output", "conv_fp32_kernel"]
cur, funcs = None, collections.OrderedDict()
for line in open(path, errors="replace"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
print("%-78s %6s " % ("kernel (mangled)", "instr") + " ".join("%7s" % k for k in KEYS))
for name, lines in funcs.items():
    ops = collections.Counter()
    n = 0
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k) or (k == "2CTA" and ".2CTA" in op):
                ops[k] += 1
    print("%-78s %6d " % (name[:78], n) + " ".join("%7d" % ops[k] for k in KEYS))
if out_dir:
    with gzip.open(out_dir + "/r2_sass_default_path_kernels.txt.gz", "wt") as f:
        for name, lines in funcs.items():
            if any(d in name for d in DEFAULT):
                f.write("Function : %s\n" % name)
                f.writelines(lines)
