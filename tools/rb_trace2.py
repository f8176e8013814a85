import sys
import numpy as np
for f in sys.argv[1:]:
    t = np.fromfile(f, dtype=np.uint64).reshape(8, 128, 4).astype(np.int64)
    t0 = t[t > 0].min()
    print(f)
    for lt in range(12, 16):
        e = t[2, lt] - t0; d = t[6, lt] - t0
        print(lt, "begin", e[0], "tm_full seen +", e[1]-e[0], "| before ldwait +", d[0]-e[1], "after ldwait +", d[1]-d[0], "math+sts +", d[2]-d[1], "fence +", d[3]-d[2], "rest(to tile done) +", e[2]-d[3])
    for g in range(3):
        print(" stencil g",g,[tuple((t[3+g,i]-t0).tolist()) for i in range(6,9)])
    print(" mma", [tuple((t[1,i,:3]-t0).tolist()) for i in range(20,26)])
    print(" producer", [tuple((t[0,i,:2]-t0).tolist()) for i in range(20,26)])
