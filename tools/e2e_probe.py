"""Where the end-to-end step's extra time over the resident step goes (one box, one run):
    python tools/e2e_probe.py
times the headline workload as: resident replay; + loss read-back (lagged / immediate); + input copy on the main stream;
+ input copy on the copy stream (prefetch); the full pipelined and unpipelined end-to-end steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = bench.Workload("cfg2_14m_32_bf16", args, 0, 1, dev, use_graph=True, warm=3)
    r, xh, yh = w.runner, w.x_host, w.y_host
    for _ in range(3):
        r()
    torch.cuda.synchronize()

    def lagged_loss():
        r()
        r.loss_to_host_async()
        return r.previous_loss()

    def main_stream_copy():
        r(xh, yh)

    def prefetch_only():
        r.prefetch(xh, yh)
        r(staged=True)

    def copy_stream_idle():  # the H2D runs, nobody consumes it: does a concurrent DMA slow the step?
        r.prefetch(xh, yh)
        r()

    variants = [("resident", lambda: r()), ("resident + lagged loss", lagged_loss),
                ("resident + immediate loss", lambda: float(r())), ("+ H2D on the main stream", main_stream_copy),
                ("+ H2D on the copy stream, consumed", prefetch_only), ("+ H2D on the copy stream, not consumed", copy_stream_idle),
                ("e2e pipelined", w.step_host), ("e2e unpipelined", w.step_host_sync)]
    for rep in range(2):
        for name, fn in variants:
            fn()
            ms = bench.timed(fn, 10, 1, dev)
            print(f"{name:42s} {ms:8.3f} ms")


if __name__ == "__main__":
    main()
