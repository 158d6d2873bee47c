"""A NaN or inf score must turn the batch loss into NaN (as in the reference), and the next clean call must be unaffected."""
import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/nn-with-pytorch-personalized-losses_b200"]
import torch
from ltr_b200.losses.approxNDCG import approxNDCGLoss
g = torch.Generator().manual_seed(1)
s = torch.randn(256, 128, generator=g); y = torch.randint(0, 5, (256, 128), generator=g).float()
print("clean ->", float(approxNDCGLoss(s.cuda(), y.cuda(), alpha=10.0)))
s2 = s.clone(); s2[7, 5] = float("nan")
print("nan   ->", float(approxNDCGLoss(s2.cuda(), y.cuda(), alpha=10.0)))
s3 = s.clone(); s3[7, 5] = float("inf")
print("inf   ->", float(approxNDCGLoss(s3.cuda(), y.cuda(), alpha=10.0)))
print("clean again ->", float(approxNDCGLoss(s.cuda(), y.cuda(), alpha=10.0)))
