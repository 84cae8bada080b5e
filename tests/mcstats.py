"""Monte-Carlo interval arithmetic shared by the curve tests (test infrastructure).

Two independent experiments measure the same decoder (ours on Philox noise, the live reference on mt19937 noise).
Under that null hypothesis
  * block errors are iid Bernoulli per frame  ->  |p1 - p2| <= z * sqrt(p (1-p) (1/n1 + 1/n2)),  p pooled;
  * a frame's bit-error FRACTION e_f (errors / K) is iid per frame with some variance v (bit errors inside a frame
    are correlated, so v is NOT ber (1-ber) / K; it is estimated from a sample of per-frame counts)
                                              ->  |ber1 - ber2| <= z * sqrt(v (1/n1 + 1/n2)).
z is the two-sided normal quantile of a FAMILY-WISE 95 % level over the m comparisons one test makes
(Bonferroni: per-comparison level 1 - 0.05/m), so a test asserting m intervals at once still fails a correct
implementation in at most 5 % of seeds; m and z are printed by the tests."""
import math

from scipy.stats import norm


def z_familywise(m, alpha=0.05):
    return float(norm.ppf(1.0 - alpha / (2.0 * max(int(m), 1))))


def bler_halfwidth(p1, n1, p2, n2, z):
    p = (p1 * n1 + p2 * n2) / float(n1 + n2)
    return z * math.sqrt(max(p * (1.0 - p), 0.5 / (n1 + n2)) * (1.0 / n1 + 1.0 / n2))


def ber_halfwidth(frame_var, n1, n2, z):
    return z * math.sqrt(frame_var * (1.0 / n1 + 1.0 / n2))


def frame_fraction_var(msg, dec):
    """Unbiased variance of the per-frame bit-error fraction; msg, dec: [B,K] torch tensors (0 in dec = error)."""
    e = (msg.round() != dec.round()).float().mean(dim=1).double()
    return float(e.var(unbiased=True).item())
