import sys
import numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
bad = 0
for k in a.files:
    if not np.array_equal(a[k], b[k]):
        bad += 1
        print("%s differs: %d elements" % (k, (a[k] != b[k]).sum()))
print("IDENTICAL" if bad == 0 else "DIFFERENT (%d arrays)" % bad)
