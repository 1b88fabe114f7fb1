#!/usr/bin/env python
"""Which role bounds the implicit-GEMM kernel on a given layer?  Times each case with parts of the kernel switched off through
FBANET_TC_DEBUG (1 = epilogue hands the accumulator straight back, 2 = no tcgen05.mma issued, 4 = A producer does not load):
the time that does NOT drop when a role is removed is the time of the roles that remain."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import prof_conv  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="body3x3_64_64,body3x3_64_64_res,faf3x3_128_128_160,faf3x3_256_256_80,proj3x3_512_256,tail3x3_64_256_320,final3x3_64_16_640,fc1_128_512,qkv_128_384,fuse_896_64")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--modes", default="0,1,2,4,3,5,6,7")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for n in a.case.split(","):
        for m in a.modes.split(","):
            os.environ["FBANET_TC_DEBUG"] = m
            print(f"debug={m} ", end="")
            prof_conv.run_case(n, a.reps, dev)
    os.environ["FBANET_TC_DEBUG"] = "0"
