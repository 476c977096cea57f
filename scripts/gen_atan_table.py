"""Prints the diamond-angle arctangent table of auromat_b200/csrc/amt_fastmath.cuh.

theta_i = atan2(i, 64 - i) in degrees (i = 0..64) at index i, 180 - theta_i at index 128 + i (63 entries
of padding between: the sign bit of x, shifted, is the offset of the second half): entry [k*128 + i]
is the angle of the direction (+-(64 - i), i), k = 1 for negative x.  Correctly rounded (mpmath,
40 digits)."""
import mpmath

mpmath.mp.dps = 40
vals = [mpmath.degrees(mpmath.atan2(i, 64 - i)) for i in range(65)]
vals = vals + [mpmath.mpf(0)] * 63 + [180 - v for v in vals]
out = []
for j in range(0, len(vals), 3):
    out.append("    " + ", ".join(repr(float(v)) for v in vals[j:j + 3]) + ",")
print("\n".join(out))
