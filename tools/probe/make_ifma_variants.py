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
print("variants:", sorted(os.listdir(out)))
