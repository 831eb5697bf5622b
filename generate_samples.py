#!/usr/bin/env python
"""Same CLI as the reference's generate_samples.py (flags -gpu -dataset -texture -ckpt_path -seq_length -bs),
running the B200-native Model.  Start frames are read from ./assets/GT_samples/<dataset>[/<texture>]/ and the
animation is written to ./assets/results/<...>/results.gif.  Resize / normalise of the start frames and the
denorm -> uint8 GIF canvas run on the GPU (image2video_synthesis_using_cinns_b200.cli)."""
import argparse
import math
import os


def parse(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-gpu', type=str, required=True, help="Define GPU on which to run")
    parser.add_argument('-dataset', type=str, required=True, help='Specify dataset')
    parser.add_argument('-texture', type=str, help='Specify texture when using DTDB')
    parser.add_argument('-ckpt_path', type=str, required=False, help='If ckpt outside of repo')
    parser.add_argument('-seq_length', type=int, default=16)
    parser.add_argument('-bs', type=int, default=6, help='Batchsize')
    parser.add_argument('-img_path', type=str, required=False, help='Folder with start frames (default ./assets/GT_samples/...)')
    parser.add_argument('-save_path', type=str, required=False, help='Output folder (default ./assets/results/...)')
    return parser.parse_args(argv)


def main(argv=None):
    args = parse(argv)
    os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu      # before the first CUDA call, like the reference

    import torch

    from get_model import Model
    from image2video_synthesis_using_cinns_b200 import cli

    path_ds = f'{args.dataset}/{args.texture}' if args.dataset == 'DTDB' else f'{args.dataset}'
    ckpt_path = f'./models/{path_ds}/stage2/' if not args.ckpt_path else args.ckpt_path
    img_path = args.img_path or f'./assets/GT_samples/{path_ds}/'

    model = Model(ckpt_path, args.seq_length)
    img_res = model.config.Data['img_size']
    names = cli.list_images(img_path)
    if not names:
        raise SystemExit(f'no start frames (*.jpg, *.png, *.jpeg) under {img_path}')
    imgs = cli.load_images(names, img_res, model.device)

    videos = []
    with torch.no_grad():
        for i in range(math.ceil(imgs.size(0) / args.bs)):
            videos.append(model(imgs[i * args.bs:(i + 1) * args.bs]))
    videos = torch.cat(videos)

    save_path = args.save_path or f'./assets/results/{path_ds}/'
    cli.save_gif(os.path.join(save_path, 'results.gif'), cli.convert_seq2gif_u8(videos), fps=3)
    print(f'Animations saved in {save_path}')
    return videos


if __name__ == "__main__":
    main()
