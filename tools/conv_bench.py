"""3x3 conv microbench (halo-tile scheme where eligible): python tools/conv_bench.py"""
import sys
sys.path.insert(0, 'tools')
import gemm_bench as g

g.bench_conv(64, 64, 320, 320, 160)
g.bench_conv(64, 64, 640, 320, 160)
g.bench_conv(64, 64, 640, 320, 224)
g.bench_conv(64, 64, 960, 320, 160)
g.bench_conv(64, 32, 640, 640, 160)
g.bench_conv(64, 32, 640, 640, 224)
g.bench_conv(64, 16, 1280, 1280, 256)
g.K.check_abort()
