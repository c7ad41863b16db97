import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv
tile = int(sys.argv[1]); G = int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 9
env = BatchedGoEnv(G, board_size=n, tile=tile)
for _ in range(3):
    env.reset(); env.playout(env.max_step + 1, seed=1)
torch.cuda.synchronize()
