"""Which hardware scoreboard does every variable-latency instruction of a kernel signal, and which instructions wait on it?
Decodes the control fields of the 128-bit sm_100 instruction words printed by `cuobjdump -sass` (bits 105..121 of each
instruction: stall count, yield, write barrier 110-112, read barrier 113-115, wait mask 116-121).  No GPU needed.

    python tools/sass_scoreboards.py multimodal-fusion-fpn_b200/csrc/conv_ws.o ILi0ELi1ELb0E [regex-of-instructions-to-list]

prints, in program order, the instructions matching the regex (default: the epilogue's 256-bit global loads / stores and the
TMEM loads) with their barriers, and every instruction whose wait mask includes a barrier that one of the listed loads signals.
This is how DESIGN.md section 5.1 found that all addend loads of the conv epilogue share one scoreboard."""
import re
import subprocess
import sys

obj, func = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else r'^(LDG\.E\.\S*256|STG\.E\.\S*256|LDTM)')
text = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
on, lines = False, []
for ln in text.split('\n'):
    if 'Function :' in ln:
        on = func in ln
    elif on:
        lines.append(ln)
ins, i = [], 0
while i < len(lines) - 1:
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', lines[i])
    m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1]) if m else None
    if m and m2:
        hi = int(m2.group(1), 16)
        ins.append((m.group(1), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
        i += 2
    else:
        i += 1
if not ins:
    sys.exit(f'no instructions found for a function matching {func!r} in {obj}')
load_bars = {wb for _, t, _, wb, _, _ in ins if pat.match(t) and t.startswith('LDG') and wb < 6}
print(f'{len(ins)} instructions; barriers signalled by the listed global loads: {sorted(load_bars)}')
started = False
for pc, t, stall, wb, rb, wm in ins:
    if pat.match(t):
        started = True
        print(f'{pc}  {t[:64]:64s} write-barrier {wb if wb < 6 else "-"}  read-barrier {rb if rb < 6 else "-"}')
    elif started and any(wm & (1 << b) for b in load_bars):
        print(f'{pc}      waits {wm:06b}: {t[:60]}')
