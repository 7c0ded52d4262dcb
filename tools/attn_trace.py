"""Timestamps (globaltimer, ns) of one CTA of attention_tc built with -DLVCB200_FA_TRACE (lvc_b200/_build/libfatrace.so): per key block j,
softmax warp 2: 0 before wait S, 1 S ready, 2 pass 1 done, 3 row max agreed, 4 P written + arrive, 5 O ready, 6 O accumulated;
MMA thread: 8 before wait K, 9 K ready, 10 S issued, 11 P ready, 12 V / O free, 13 P V issued."""
import ctypes, sys
import torch
lib = ctypes.CDLL("lvc_b200/_build/libfatrace.so")
B, N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 42, 896, 1
qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty((B * N, H * 64), dtype=torch.bfloat16, device="cuda")
lib.lvcb200_attention_tc.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
for _ in range(3):
    rc = lib.lvcb200_attention_tc(qkv.data_ptr(), B, N, H, 64, ctypes.c_float(0.125), out.data_ptr(), None)
    assert rc == 0
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 256)()
lib.lvcb200_debug_attention_trace(buf)
t0 = buf[0]
for j in range(7):
    r = [buf[16 * j + k] - t0 for k in range(14)]
    print(f"blk {j}: softmax wait_s {r[0]:6d} s_ready {r[1]:6d} pass1 {r[2]:6d} max {r[3]:6d} p_done {r[4]:6d} o_ready {r[5]:6d} o_acc {r[6]:6d} | "
          f"mma wait_k {r[8]:6d} k_ready {r[9]:6d} wait_p {r[10]:6d} p_ready {r[11]:6d} v_ready {r[12]:6d} pv_issued {r[13]:6d}")
