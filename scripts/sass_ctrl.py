"""Annotates `cuobjdump -sass` output with the scheduling control fields of every instruction (sm_70+ encoding:
stall count bits [105,109), yield bit 109, write / read scoreboard slots, wait mask bits [116,122)), so a kernel's hot
loop can be checked for exposed latencies here, without a GPU.
Usage: cuobjdump -sass -fun <mangled> file.o | python scripts/sass_ctrl.py [first_addr last_addr]"""
import re
import sys

lo, hi = (int(sys.argv[1], 16), int(sys.argv[2], 16)) if len(sys.argv) > 2 else (0, 1 << 62)
lines = sys.stdin.read().split("\n")
pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
hipat = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
for i, line in enumerate(lines):
    m = pat.search(line)
    if not m:
        continue
    addr = int(m.group(1), 16)
    if addr < lo or addr > hi:
        continue
    h = hipat.match(lines[i + 1]) if i + 1 < len(lines) else None
    if not h:
        continue
    w = int(h.group(1), 16)
    stall = (w >> 41) & 0xF
    yld = (w >> 45) & 1
    wbar = (w >> 46) & 7
    rbar = (w >> 49) & 7
    wait = (w >> 52) & 0x3F
    print(f"{addr:05x} s{stall:<2d} {'Y' if yld else ' '} w{wbar if wbar != 7 else '-'} r{rbar if rbar != 7 else '-'} "
          f"wait{wait:06b}  {m.group(2).strip()}")
