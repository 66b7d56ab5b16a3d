import sys, os
sys.argv = ["x", "none"]
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import probe_kernels as P
P.fft_probe((416, 16 * 416 * 416))
P.fft_probe((416, 416, 416 * 16))
P.fft_probe((416, 416, 416, 16))
P.fft_probe((512, 512, 256, 12))
