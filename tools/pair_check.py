"""CTA-pair (cta_group::2) variants of the padded-flat conv kernel against the single-CTA kernel: outputs must be bit-identical
(same accumulation order), the deferred per-channel sums equal to ~1e-6. Shapes forced through CILRS_FLAT_SHAPE."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops

torch.manual_seed(0)
bad = 0
for B in (128, 7):
    for h, w, c in [(22, 50, 64), (11, 25, 128), (6, 13, 256), (3, 7, 512)]:
        d = ops.conv_desc(B, h, w, c, c, 3, 1)
        x = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
        wf, wd = ops.pack_weight(d, torch.randn(c, c, 3, 3, device="cuda") * 0.05)
        act = ops.to_padded(torch.relu(torch.randn(B, h, w, c, device="cuda")).to(torch.bfloat16))
        y1 = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
        res = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
        bits = ops.relu_bits(act)

        def run():
            a = ops.conv_flat(x, wf, c)
            b, sb = ops.conv_flat(x, wf, c, defer_sums="stats")
            cc, sc = ops.conv_flat(x, wd, c, dgrad=True, residual=res, mask=act, mask_bits=bits, bnbwd=dict(y=y1), defer_sums="bnbwd")
            torch.cuda.synchronize()
            return a, b, sb, cc, sc

        os.environ["CILRS_FLAT_SHAPE"] = "1,64,0,3,0"
        ref = run()
        for mt in (1, 2, 4):
            for bn in (64, 128, 256):
                if c % bn or mt * bn > 512:
                    continue
                for r, G in ((1, 9), (0, 3), (0, 1)):
                    os.environ["CILRS_FLAT_SHAPE"] = "%d,%d,%d,%d,1" % (mt, bn, r, G)
                    try:
                        got = run()
                    except RuntimeError as e:
                        continue
                    ok = torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and torch.equal(got[3], ref[3])
                    e1 = ((got[2] - ref[2]).abs().max() / ref[2].abs().max()).item()
                    e2 = ((got[4] - ref[4]).abs().max() / ref[4].abs().max()).item()
                    ok = ok and e1 < 1e-5 and e2 < 1e-5
                    bad += 0 if ok else 1
                    print("B=%d c=%d pair mt%d bn%d r%d G%d: %s (sums %.1e %.1e)" % (B, c, mt, bn, r, G, "OK" if ok else "MISMATCH", e1, e2))
print("mismatches:", bad)
sys.exit(1 if bad else 0)
