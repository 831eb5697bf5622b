#!/usr/bin/env python
"""Same CLI as the reference's generate_transfer.py (flags -gpu -dataset -ckpt_path -seq_length): every clip
under ./assets/GT_samples/<dataset>/transfer/ is used in turn as the motion query for all first frames."""
import argparse
import math
import os


def parse(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-gpu', type=str, required=True, help="Define GPU on which to run")
    parser.add_argument('-dataset', type=str, required=True, help='Specify dataset')
    parser.add_argument('-ckpt_path', type=str, required=False, help='If ckpt outside of repo')
    parser.add_argument('-seq_length', type=int, default=16)
    parser.add_argument('-img_path', type=str, required=False, help='Folder with one sub-folder of frames per clip')
    parser.add_argument('-save_path', type=str, required=False, help='Output folder (default ./assets/results/...)')
    return parser.parse_args(argv)


def main(argv=None):
    args = parse(argv)
    os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu

    import torch

    from get_model import Model
    from image2video_synthesis_using_cinns_b200 import cli

    ckpt_path = f'./models/{args.dataset}/stage2/' if not args.ckpt_path else args.ckpt_path
    model = Model(ckpt_path, args.seq_length, transfer=True)
    img_path = args.img_path or f'./assets/GT_samples/{args.dataset}/transfer/'
    img_res = model.config.Data['img_size']

    videos = []
    for vid in sorted(os.listdir(img_path), key=cli.natural_key):
        frames = sorted(cli.list_images(os.path.join(img_path, vid)), key=cli.natural_key)[:args.seq_length]
        if frames:
            videos.append(cli.load_images(frames, img_res, model.device))
    if not videos:
        raise SystemExit(f'no clips under {img_path}')
    videos = torch.stack(videos)

    bs = 6
    save_path = args.save_path or f'./assets/results/{args.dataset}/'
    results = []
    for idx, query in enumerate(videos):
        transfer = []
        with torch.no_grad():
            for i in range(math.ceil(videos.size(0) / bs)):
                transfer.append(model.transfer(query[None], videos[i * bs:(i + 1) * bs, 0]))
        transfer = torch.cat((query[None, :transfer[0].shape[1]], torch.cat(transfer)), dim=0)
        cli.save_gif(os.path.join(save_path, f'transfer_{idx}.gif'), cli.convert_seq2gif_u8(transfer), fps=3)
        results.append(transfer)
    print(f'Transfers saved in {save_path}')
    return results


if __name__ == "__main__":
    main()
