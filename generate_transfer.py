#!/usr/bin/env python
"""Same CLI as the reference's generate_transfer.py (flags -gpu -dataset -ckpt_path -seq_length): every clip
under ./assets/GT_samples/<dataset>/transfer/ is used in turn as the motion query for all first frames."""
import argparse
import math
import os

parser = argparse.ArgumentParser()
parser.add_argument('-gpu', type=str, required=True, help="Define GPU on which to run")
parser.add_argument('-dataset', type=str, required=True, help='Specify dataset')
parser.add_argument('-ckpt_path', type=str, required=False, help='If ckpt outside of repo')
parser.add_argument('-seq_length', type=int, default=16)
parser.add_argument('-img_path', type=str, required=False, help='Folder with one sub-folder of frames per clip')
args = parser.parse_args()
os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu

import torch  # noqa: E402

from get_model import Model  # noqa: E402
from image2video_synthesis_using_cinns_b200 import cli  # noqa: E402

ckpt_path = f'./models/{args.dataset}/stage2/' if not args.ckpt_path else args.ckpt_path
model = Model(ckpt_path, args.seq_length, transfer=True)
img_path = args.img_path or f'./assets/GT_samples/{args.dataset}/transfer/'
img_res = model.config.Data['img_size']

videos = []
for vid in sorted(os.listdir(img_path), key=cli.natural_key):
    frames = sorted(cli.list_images(os.path.join(img_path, vid)), key=cli.natural_key)[:args.seq_length]
    if frames:
        videos.append(torch.stack([cli.load_image(n, img_res) for n in frames]))
if not videos:
    raise SystemExit(f'no clips under {img_path}')
videos = torch.stack(videos)

bs = 6
save_path = f'./assets/results/{args.dataset}/'
for idx, query in enumerate(videos):
    transfer = []
    with torch.no_grad():
        for i in range(math.ceil(videos.size(0) / bs)):
            transfer.append(model.transfer(query[None].cuda(), videos[i * bs:(i + 1) * bs, 0].cuda()).cpu())
    transfer = torch.cat((query[None, :transfer[0].shape[1]], torch.cat(transfer)), dim=0)
    cli.save_gif(save_path + f'transfer_{idx}.gif', cli.convert_seq2gif(transfer), fps=3)
print(f'Transfers saved in {save_path}')
