"""Timing-only ablations / alternatives of the IFMA Poseidon path for tools/probe/run_poseidon_lab.sh: copies of
sipp_b200/csrc/poseidon_avx512.cc with pieces of the partial-round loop removed (wrong results -- the lab only times them).
Writes tools/probe/variants/i_*.cc."""
import os
import sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(root, "sipp_b200/csrc/poseidon_avx512.cc")).read()
out = os.path.join(root, "tools/probe/variants")
os.makedirs(out, exist_ok=True)


def cut(s, start, end):
    a = s.index(start)
    b = s.index(end, a)
    return s[:a] + s[b:]


def write(name, s):
    open(os.path.join(out, name), "w").write(s)


upd_start = "        if (j >= 1) {  // the vector terms of x_{j-1}"
upd_end = "        if (j + 1 <= 21) {  // close C-row j + 1"
close_line = "            e = row_close(a0, a1, a2, I.cdiag[row], p7);\n"
novec = cut(src, upd_start, upd_end)
noclose = src.replace(close_line, "")
write("i_novec.cc", novec)
write("i_noclose.cc", noclose)
write("i_neither.cc", cut(noclose, upd_start, upd_end))
write("i_mds_bb.cc", "#define LAB_MDS_BB 1\n" + src)            # rows 8..11 of the MDS layer from a merged (low | high) broadcast, closed on the vector ports
write("i_oldmul.cc", "#define LAB_OLD_MODMUL 1\n" + src)       # the 25-micro-op vector product on the one-vector chain
write("i_mds_bb_oldmul.cc", "#define LAB_MDS_BB 1\n#define LAB_OLD_MODMUL 1\n" + src)
print("variants:", sorted(os.listdir(out)))
