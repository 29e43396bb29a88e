import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_wide
bench_wide.svgd = lambda *a, **k: None
bench_wide.predictive()
