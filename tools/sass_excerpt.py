"""Cut the SASS of a source-line range out of a kernel: `nvdisasm -g -c` keeps '//## File ..., line N' markers
(build with -lineinfo).  usage: sass_excerpt.py <file.cu> <kernel substring> <first line> <last line>
Extracts the cubin from jdeflate_b200/lib/libjdeflate.so into a temporary directory, prints the instructions
attributed to those lines with their source line in front, and a count by mnemonic at the end."""
import collections, pathlib, re, subprocess, sys, tempfile
cu, kern, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
root = pathlib.Path(__file__).resolve().parent.parent
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "jdeflate_b200/lib/libjdeflate.so")], cwd=tmp, check=True, capture_output=True)
cubin = next(tmp.glob(cu.split(".")[0] + "*.cubin"))
dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True, check=True).stdout
cur = None; line = None; n = 0; ops = collections.Counter()
src = (root / "jdeflate_b200/csrc/device" / cu).read_text().split("\n")
last_printed = None
for l in dis.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", l)
    if m and cur and kern in cur and line and line[0] == cu and a <= line[1] <= b:
        if last_printed != line[1]:
            print("%s:%d  %s" % (cu, line[1], src[line[1] - 1].strip()[:110]))
            last_printed = line[1]
        ins = m.group(2)
        print("        /*%s*/  %s" % (m.group(1), ins))
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
        ops[op.split(".")[0]] += 1; n += 1
print("\n%d instructions; by mnemonic: %s" % (n, ", ".join("%s %d" % kv for kv in ops.most_common())))
