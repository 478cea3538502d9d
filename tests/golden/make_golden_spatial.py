"""Golden boxes of the REAL reference sliding_window (ever/magic/bigimage/sliding_window.py), build container only:

    python tests/golden/make_golden_spatial.py        # writes tests/golden/sliding_window_boxes.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, '_stubs'))
from ever.magic.bigimage.sliding_window import sliding_window  # noqa: E402

CASES = [((1024, 1024), 512, 256), ((1000, 1500), 512, 384), ((300, 700), 512, 256), ((513, 512), (512, 256), (128, 256)),
         ((2048, 1024), (512, 512), (512, 512)), ((97, 33), 32, 7)]

if __name__ == '__main__':
    out = [dict(input_size=size, kernel_size=k, stride=s, boxes=sliding_window(size, k, s).tolist()) for size, k, s in CASES]
    json.dump(out, open(os.path.join(HERE, 'sliding_window_boxes.json'), 'w'))
    print(sum(len(o['boxes']) for o in out), 'boxes')
