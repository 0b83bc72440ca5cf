import os, sys, torch
sys.path.insert(0, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200")
from b200 import engine
dev = torch.device("cuda:0")
try:
    x = torch.zeros(1, 16000, device=dev)
    y = engine.logmel(x)
    torch.cuda.synchronize()
    print("ok", y.shape, y[0, 0, :4])
except Exception as e:
    print("ERR", repr(e)[:600])
