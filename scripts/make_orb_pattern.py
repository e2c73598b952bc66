"""Transcribes the 256 learned rBRIEF test pairs of ORB (Rublee et al., "ORB: an efficient alternative to SIFT or SURF", 2011;
OpenCV's orb.cpp publishes them as bit_pattern_31_, and the reference carries that table at [FEAT]:448-705) into a bare
number table: 256 rows of x0, y0, x1, y1.  The numbers are data of the published algorithm -- descriptors are only
comparable between implementations that use the same pairs -- not code.  Run where /root/reference is present:

    python scripts/make_orb_pattern.py     # writes oracle/orb_pattern.inc and imagestitch_b200/csrc/orb_pattern.inc
"""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = [p for p in glob.glob("/root/reference/*/*/*.cpp") if b"bit_pattern_31_" in open(p, "rb").read()]
    assert src, "reference not present"
    text = open(src[0], "rb").read().decode("latin-1")
    body = text[text.index("bit_pattern_31_[256 * 4]"):]
    body = body[body.index("{") + 1:body.index("};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    nums = [int(v) for v in re.findall(r"-?\d+", body)]
    assert len(nums) == 1024 and all(-15 <= v <= 15 for v in nums)
    lines = ["// rBRIEF test pairs of ORB (published with the algorithm; OpenCV orb.cpp bit_pattern_31_): x0, y0, x1, y1 per row.",
             "// Written by scripts/make_orb_pattern.py -- data, do not edit."]
    for i in range(0, 1024, 16):
        lines.append(" ".join(f"{v}," for v in nums[i:i + 16]))
    out = "\n".join(lines) + "\n"
    for rel in ("oracle/orb_pattern.inc", "imagestitch_b200/csrc/orb_pattern.inc"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(out)
    print("wrote 256 pairs")


if __name__ == "__main__":
    main()
