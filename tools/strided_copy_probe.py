"""Can the copy engines skip the zero (strictly upper) part of L and the unread upper part of A?
Measures pinned host <-> device copies of column blocks of a batch of n x n f64 matrices with
cudaMemcpy3DAsync (row width w bytes, row pitch 8 n, matrix pitch 8 n n), against the contiguous
copy of the whole batch -- the evidence for/against a "block triangle" host path.

    python tools/strided_copy_probe.py [n] [batch]
"""
import sys
import time

import torch
from cuda.bindings import runtime as rt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
nbytes = batch * n * n * 8
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d2 = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost


def copy3d(dst, src, kind, stream, row0, nrows, col0_bytes, w_bytes):
    """rows row0..row0+nrows, bytes col0..col0+w of every matrix"""
    p = rt.cudaMemcpy3DParms()
    off = row0 * n * 8 + col0_bytes
    p.srcPtr = rt.make_cudaPitchedPtr(src + off, n * 8, n * 8, n)
    p.dstPtr = rt.make_cudaPitchedPtr(dst + off, n * 8, n * 8, n)
    p.extent = rt.make_cudaExtent(w_bytes, nrows, batch)
    p.kind = kind
    (err,) = rt.cudaMemcpy3DAsync(p, stream.cuda_stream)
    assert err == rt.cudaError_t.cudaSuccess, err


def blocks(nb):
    """lower block triangle with nb column blocks: (row0, nrows, col0_bytes, w_bytes)"""
    w = n // nb
    return [(c * w, n - c * w, c * w * 8, w * 8) for c in range(nb)]


def run(label, nb, dirs, reps=3):
    bl = blocks(nb)
    moved = sum(b[1] * b[3] for b in bl) * batch
    for it in range(reps + 1):
        if it == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        for b in bl:
            if "h2d" in dirs:
                if nb == 1:
                    (err,) = rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), nbytes, H2D, s_in.cuda_stream)
                else:
                    copy3d(d.data_ptr(), h.data_ptr(), H2D, s_in, *b)
            if "d2h" in dirs:
                if nb == 1:
                    (err,) = rt.cudaMemcpyAsync(h2.data_ptr(), d2.data_ptr(), nbytes, D2H, s_out.cuda_stream)
                else:
                    copy3d(h2.data_ptr(), d2.data_ptr(), D2H, s_out, *b)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    nd = len(dirs)
    print(f"n={n} {label:28s} {'+'.join(dirs):8s} payload {moved / nbytes * 100:5.1f}% of dense  "
          f"{nd * moved / dt / 1e9:7.1f} GB/s moved   {batch / dt / 1e6:7.2f} M matrices/s", flush=True)


for dirs in (("h2d",), ("d2h",), ("h2d", "d2h")):
    run("dense (one memcpy)", 1, dirs)
    for nb in (2, 4, 8):
        if n % nb == 0 and (n // nb) * 8 >= 32:
            run(f"{nb} column blocks (w={n // nb * 8} B)", nb, dirs)
# correctness of the addressing: the skipped part of the destination is untouched
h2.zero_()
for b in blocks(2):
    copy3d(h2.data_ptr(), d2.data_ptr(), D2H, s_out, *b)
torch.cuda.synchronize()
m = h2.view(torch.float64).view(batch, n, n)
one = torch.ones(1, dtype=torch.uint8).repeat(8).view(torch.float64).item()
assert (m[:, :, : n // 2] == one).all() and (m[:, n // 2:, n // 2:] == one).all() and (m[:, : n // 2, n // 2:] == 0).all()
print("block addressing ok")
