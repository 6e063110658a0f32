"""Raw host<->device copy rates of the box, to put the end-to-end (host-array) numbers of bench.py in context:
pinned H2D alone, D2H alone, both directions at once, and H2D in staging-sized pieces (the library moves
16 384-column pieces, ~9.4 MB per (ncol, nlay) array).  Prints one JSON line."""
import json
import time

import torch


def rate(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(); torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return round(best, 2)


def main():
    n = 1 << 28                                              # 2 GiB of fp64
    h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_in.fill_(1.0)
    h_out = torch.empty(n // 4, dtype=torch.float64).pin_memory()
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.ones(n // 4, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    piece = 16384 * 72                                       # one staged (ncol, nlay) array

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    def h2d_pieces():
        with torch.cuda.stream(s1):
            for o in range(0, n - piece + 1, piece):
                d_in[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)

    out = {"h2d_gbs": rate(h2d, n * 8), "d2h_gbs": rate(d2h, n * 2),
           "h2d_while_d2h_gbs_total": rate(both, n * 8 + n * 2),
           "h2d_9MB_pieces_gbs": rate(h2d_pieces, (n // piece) * piece * 8),
           "gpu": torch.cuda.get_device_name(0)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
