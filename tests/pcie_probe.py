"""B200 box: host<->device copy rates for the sizes of the bench's e2e step (not a pytest test).

    python tests/pcie_probe.py

Prints GB/s of pinned H2D / D2H copies alone and concurrently (two streams), per size: the floor under the
pipelined host-data path (bsmr_sddmm_host_submit) is max(H2D bytes / H2D rate, D2H bytes / D2H rate)."""
import json
import time

import torch


def rate(fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    for mb in (0.77, 3.0, 6.4, 64.0):
        n = int(mb * 1e6 / 4)
        h_in = torch.empty(n, dtype=torch.float32).pin_memory()
        h_out = torch.empty(n, dtype=torch.float32).pin_memory()
        d_in = torch.empty(n, dtype=torch.float32, device="cuda")
        d_out = torch.empty(n, dtype=torch.float32, device="cuda")

        def h2d():
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)

        def both():
            h2d()
            d2h()

        def h2d_split():     # the same bytes as two halves on two streams (two copy engines)
            half = n // 2
            with torch.cuda.stream(s_in):
                d_in[:half].copy_(h_in[:half], non_blocking=True)
            with torch.cuda.stream(s_out):
                d_in[half:].copy_(h_in[half:], non_blocking=True)

        print(json.dumps({"MB": mb, "h2d_GBs": rate(h2d, n * 4), "d2h_GBs": rate(d2h, n * 4),
                          "both_each_GBs": rate(both, n * 4), "h2d_two_streams_GBs": rate(h2d_split, n * 4)}), flush=True)


if __name__ == "__main__":
    main()
