"""Training sub-benchmark alone (BASELINE config #3), repeated: python scripts/train_bench.py [reps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    info = bench.run_train(0, 1, dev, 50, 5)
    print(f"rep {r}: {info['value']:.1f} it/s  {info['ms_per_iter']:.3f} ms/iter  loss {info['loss']:.4f}")
