"""Is the pitched (2-D) device-to-host copy of the host-buffer Gram leg slower than a flat one? (development aid)"""
import ctypes as C, time, torch
rt = C.CDLL("libcudart.so")
rows, w_cols, ldk = 2048, 49152, 65536           # one 2048-row block of the N = 65536 trapezoid, 3/4 down
dev = torch.empty(rows * w_cols, dtype=torch.float64, device="cuda").normal_()
host = torch.empty(rows * ldk, dtype=torch.float64).pin_memory(); host.zero_()
flat = torch.empty(rows * w_cols, dtype=torch.float64).pin_memory(); flat.zero_()
s = torch.cuda.current_stream().cuda_stream
def t(fn, reps=8):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return rows * w_cols * 8 * reps / (time.perf_counter() - t0) / 1e9
def copy2d(h_ptr, pitch, d_ptr, dpitch, width, height):
    rc = rt.cudaMemcpy2DAsync(C.c_void_p(h_ptr), C.c_size_t(pitch), C.c_void_p(d_ptr), C.c_size_t(dpitch), C.c_size_t(width), C.c_size_t(height), C.c_int(2), C.c_void_p(s))
    assert rc == 0, rc
def copy1d(h_ptr, d_ptr, nbytes):
    rc = rt.cudaMemcpyAsync(C.c_void_p(h_ptr), C.c_void_p(d_ptr), C.c_size_t(nbytes), C.c_int(2), C.c_void_p(s))
    assert rc == 0, rc
print(f"flat 1-D copy                     : {t(lambda: copy1d(flat.data_ptr(), dev.data_ptr(), rows * w_cols * 8)):.1f} GB/s")
print(f"2-D copy, host pitch = N          : {t(lambda: copy2d(host.data_ptr(), ldk * 8, dev.data_ptr(), w_cols * 8, w_cols * 8, rows)):.1f} GB/s")
print(f"2-D copy, both pitches = width    : {t(lambda: copy2d(flat.data_ptr(), w_cols * 8, dev.data_ptr(), w_cols * 8, w_cols * 8, rows)):.1f} GB/s")
def per_row():
    for r in range(rows): copy1d(host.data_ptr() + r * ldk * 8, dev.data_ptr() + r * w_cols * 8, w_cols * 8)
print(f"one 1-D copy per row (2048 calls) : {t(per_row, reps=3):.1f} GB/s")
for wc in (2048, 8192, 16384):
    d2 = dev[: rows * wc]
    print(f"2-D copy, width {wc:5d} doubles      : {rows * wc * 8 / (rows * w_cols * 8) * t(lambda: copy2d(host.data_ptr(), ldk * 8, d2.data_ptr(), wc * 8, wc * 8, rows)):.1f} GB/s")
