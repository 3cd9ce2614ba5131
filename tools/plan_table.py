"""Launch plan of every conv layer of the codec on the plane engine (host logic only, runs without a GPU):
   python tools/plan_table.py [frames]   ->  markdown table"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nsc_b200 import _lib  # noqa: E402

KEYS = ('kind', 'staged', 'pair', 'mt', 'n_iss', 'resident', 'wslots', 'stages', 'smem', 'tmem_cols', 'grid', 'units')
LAYERS = [   # (name, Lin, Cin, Cout, k, dil, stride, res_mode, shuffle) in launch order of one codec
    ('enc stem k55 1->100', 512, 1, 100, 55, 1, 1, 0, 1),
    ('block conv 1 100->20 @512', 512, 100, 20, 9, 1, 1, 0, 1),
    ('block conv 2 20->20 d1 @512', 512, 'narrow', 1),
    ('block conv 2 20->20 d2 @512', 512, 'narrow', 2),
    ('block conv 3 20->100 + res @512', 512, 20, 100, 9, 1, 1, 1, 1),
    ('down 100->100 stride 2', 512, 100, 100, 9, 1, 2, 0, 1),
    ('block conv 1 100->20 @256', 256, 100, 20, 9, 1, 1, 0, 1),
    ('block conv 2 20->20 d1 @256', 256, 'narrow', 1),
    ('block conv 2 20->20 d2 @256', 256, 'narrow', 2),
    ('block conv 3 20->100 + res @256', 256, 20, 100, 9, 1, 1, 1, 1),
    ('code head k55 100->1', 256, 100, 1, 55, 1, 1, 0, 1),
    ('dec k9 1->20', 256, 1, 20, 9, 1, 1, 0, 1),
    ('dec block conv 3 20->100 + bcast', 256, 20, 100, 9, 1, 1, 2, 1),
    ('up 100->100 + sub-pixel', 256, 100, 100, 9, 1, 1, 0, 2),
    ('dec block conv 1 50->20', 512, 50, 20, 9, 1, 1, 0, 1),
    ('dec block conv 3 20->50 + res', 512, 20, 50, 9, 1, 1, 1, 1),
    ('out head k55 50->1', 512, 50, 1, 55, 1, 1, 0, 1),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
    lib = _lib.load()
    fam = {0: 'taps-in-N', 1: 'tap-shift', 2: 'Toeplitz'}
    print(f'| layer ({B} frames per pass) | kernel | CTA pair | M tiles | issuers | weights | input stages | smem KB | TMEM cols |')
    print('|---|---|---|---|---|---|---|---|---|')
    for layer in LAYERS:
        out = (C.c_int64 * 12)()
        name, L = layer[0], layer[1]
        if layer[2] == 'narrow':     # the block's second conv, as the codec program runs it (folded images where the frame is long enough)
            rc = lib.nsc_narrow_conv_plan_info(B, L, layer[3], out)
        else:
            _, L, cin, cout, k, dil, stride, res, sh = layer
            rc = lib.nsc_conv1d_tc_plan_info(B, L, cin, cout, k, dil, stride, res, sh, 1, out)
        assert rc == 0, _lib.last_error()
        p = dict(zip(KEYS, list(out)))
        if p['staged'] in (5, 6):
            kern = 'tap-shift on folded images, ' + ('48->48 k5' if p['staged'] == 5 else '48->48 k9 block-diagonal') + ', staged unfold'
        elif p['kind'] == 0:
            kern = 'taps-in-N' + (f" ({p['staged']} tap groups)" if p['staged'] else '')
        else:
            kern = fam[p['kind']] + (', staged epilogue' if p['staged'] else '')
        w = f"resident ({p['wslots']} units)" if p['resident'] else f"ring of {p['wslots']}"
        print(f"| {name} | {kern} | {'yes' if p['pair'] else 'no'} | {p['mt']} | {p['n_iss']} | {w} | {p['stages']} | {p['smem'] / 1024:.0f} | {p['tmem_cols']} |")


if __name__ == '__main__':
    main()
