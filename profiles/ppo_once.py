"""One PPO train() call (config 3: 4096 x 128 samples, (200, 100) tanh MLPs, 25 epochs, normalisers
on) between cudaProfilerStart / cudaProfilerStop after two warm-up calls -- target of

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --profile-from-start off -c 600 --csv --log-file ppo_launches.csv \
        python profiles/ppo_once.py [--epochs 2]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import configs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--epochs', type=int, default=2)
args = ap.parse_args()
dev = torch.device('cuda:0')
agent = configs._ppo_agent(dev, 17, 6, args.epochs, normalize=True)
exp = configs._ppo_experience(dev, 4096, 128, 17, 6, 100)
for _ in range(2):
  agent.train(exp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
info = agent.train(exp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('ok', float(info.loss.item()))
