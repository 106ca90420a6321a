"""Decode the scheduling-control bits of sm_100a SASS (cuobjdump -sass output on stdin or a file):
stall count, yield, write-barrier index, read-barrier index, wait mask -- to see WHICH scoreboard an instruction waits on
and which earlier instruction arms it.   cuobjdump -sass -fun NAME lib.so | python profiles/sassctl.py [regex-of-anchor] [before] [after]"""
import re
import sys


def parse(text):
    lines = text.split('\n')
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append(dict(addr=m.group(1), op=m.group(2).strip(), stall=(hi >> 41) & 0xf, yld=(hi >> 45) & 1,
                                wr=(hi >> 46) & 7, rd=(hi >> 49) & 7, wait=(hi >> 52) & 0x3f))
                i += 2
                continue
        i += 1
    return ins


def fmt(q):
    wr = '-' if q['wr'] == 7 else str(q['wr'])
    rd = '-' if q['rd'] == 7 else str(q['rd'])
    return f"{q['addr']} st{q['stall']:2d} y{q['yld']} W{wr} R{rd} wait{q['wait']:06b}  {q['op']}"


if __name__ == '__main__':
    ins = parse(sys.stdin.read())
    pat = sys.argv[1] if len(sys.argv) > 1 else None
    before = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    after = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    if not pat:
        for q in ins:
            print(fmt(q))
    else:
        for n, q in enumerate(ins):
            if re.search(pat, q['op']):
                for r in ins[max(0, n - before):n + after]:
                    print(fmt(r))
                print('=' * 40)
