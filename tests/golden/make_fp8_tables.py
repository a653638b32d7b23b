"""Generates the fp8(e4m3fn)<->bf16 golden tables from torch CPU (run in the build container).

    python tests/golden/make_fp8_tables.py

The reference (ai-dynamo/dynamo) has no fp8 cast on the KV copy path, so parity
for the cast is pinned against torch CPU instead: `uint8.view(float8_e4m3fn).to(bfloat16)` for all
256 codes, and `bfloat16.to(float8_e4m3fn)` for all 65 536 bf16 bit patterns.
torch's down-cast maps overflow to NaN; the transfer kernel saturates to +-448 (vLLM's fp8 KV
convention) -- tests compare against torch wherever torch's answer is not an overflow-NaN.
"""
import os
import numpy as np
import torch

here = os.path.dirname(os.path.abspath(__file__))
codes = torch.arange(256, dtype=torch.uint8)
up = codes.view(torch.float8_e4m3fn).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
np.save(os.path.join(here, "fp8_e4m3_to_bf16_torch.npy"), up)
bits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
down = bits.to(torch.float8_e4m3fn).view(torch.uint8).numpy()
np.save(os.path.join(here, "bf16_to_fp8_e4m3_torch.npy"), down)
print("torch", torch.__version__, "wrote", up.shape, down.shape)
