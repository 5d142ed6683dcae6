"""pyrDown of a 20000^2 uint16 image (and an odd-sized uint8 one): ms per call, GB/s of algorithmic traffic (2.5 B/px in).
MA_PYRDOWN_SIMPLE=1 selects the one-output-per-thread kernel for comparison."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microaligner_b200 import ops  # noqa: E402

for shape, dt in (((20000, 20000), torch.uint16), ((12501, 12503), torch.uint16), ((20001, 19999), torch.uint8)):
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.randint(0, 255, shape, device="cuda", generator=g, dtype=torch.int32).to(dt)
    for _ in range(3):
        out = ops.pyr_down(img)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        out = ops.pyr_down(img)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    px = shape[0] * shape[1]
    print(f"{'simple' if os.environ.get('MA_PYRDOWN_SIMPLE') else 'march '} {shape} {str(dt):12s} {ms:.3f} ms "
          f"{px * img.element_size() * 1.25 / ms / 1e6:.0f} GB/s  checksum {int(out.to(torch.int64).sum())}")
