"""Decode the scheduling control fields of sm_100a SASS (``cuobjdump -sass`` output): per instruction the
stall count, yield flag, the scoreboard it SETS on completion of its write (W) or of its operand read (R),
and the mask of scoreboards it WAITS for before issue.  Usage:

    cuobjdump -sass -fun '<mangled kernel>' multimodalsignal_b200/csrc/gru.o > /tmp/k.sass
    python tools/sass_ctl.py /tmp/k.sass [first_addr_hex last_addr_hex] [--no-ffma]

The control word sits in bits 41..62 of the upper 64-bit half of each 128-bit instruction
(stall 4 bits | yield 1 | write-barrier 3 | read-barrier 3 | wait mask 6 | reuse 4), the layout used since
Volta.  What it is for here: variable-latency instructions (LDG, LDS, SHFL, MUFU, S2R ...) signal completion
through one of SIX counting scoreboards per warp, and a consumer waits for the scoreboard's counter to reach
zero -- i.e. for EVERY instruction in flight on that scoreboard, not just the one that produced its operand.
`profiles/r1_gru_bwd_scoreboard.md` uses this to explain the long_scoreboard stalls of gru_bwd_kernel.
Also prints, for the address range, the instruction count and the sum of stall counts (the static issue
schedule of a single warp, before any scoreboard wait).
"""
from __future__ import annotations

import re
import sys


def decode(path, lo=0, hi=1 << 30, skip_ffma=False):
    lines = open(path).read().splitlines()
    i, n, stalls = 0, 0, 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                addr, w = int(m.group(1), 16), int(m2.group(1), 16)
                stall, yld = (w >> 41) & 0xF, (w >> 45) & 1
                wr, rd, wait = (w >> 46) & 7, (w >> 49) & 7, (w >> 52) & 0x3F
                if lo <= addr <= hi:
                    n += 1
                    stalls += stall
                    text = m.group(2).strip()
                    if not (skip_ffma and text.startswith("FFMA2")):
                        print(f"{addr:05x} stall{stall:2d} y{yld} W{wr if wr != 7 else '-'} R{rd if rd != 7 else '-'} "
                              f"wait{wait:06b}  {text[:96]}")
                i += 2
                continue
        i += 1
    print(f"# {n} instructions, sum of stall counts {stalls} cycles")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lo = int(args[1], 16) if len(args) > 1 else 0
    hi = int(args[2], 16) if len(args) > 2 else 1 << 30
    decode(args[0], lo, hi, "--no-ffma" in sys.argv)
