import sys, numpy as np
a = np.loadtxt(sys.argv[1])
edges = [0, 16, 128, 1024, 4096, 16384, 1 << 21]
print('bucket: iters | A starts, B step, C1 classify, C2 scan, total(between scans) [us, mean]')
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (a[:, 0] >= lo) & (a[:, 0] < hi) & (a[:, 5] > 0)
    if m.any():
        print('[%6d,%7d): %5d | %6.1f %7.1f %6.1f %6.1f | %7.1f' % ((lo, hi, m.sum()) + tuple(a[m, k].mean() / 1e3 for k in (1, 2, 3, 4, 5))))
