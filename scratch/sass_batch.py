"""Show, for a kernel in an object file, every run of gather loads (LDG.E.128.CONSTANT) between FFMA2 groups:
how many loads are issued back to back and whether destination registers repeat (= the batch was serialised)."""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
sass = subprocess.check_output(["cuobjdump", "-sass", obj], text=True)
blocks = re.split(r"\n\s*Function : ", sass)
for b in blocks:
    name = b.split("\n", 1)[0]
    if not re.search(pat, name):
        continue
    print("==", name[:90])
    run, regs, n_instr = 0, [], 0
    for line in b.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        n_instr += 1
        if "LDG.E.128.CONSTANT" in ins:
            run += 1
            regs.append(re.search(r"LDG\.E\.128\.CONSTANT (R\d+)", ins).group(1))
        elif "FFMA2" in ins or "FFMA " in ins:
            if run:
                print(f"  loads in flight before first FMA: {run}  dest regs {regs}  distinct {len(set(regs))}")
            run, regs = 0, []
    print("  instructions:", n_instr)
