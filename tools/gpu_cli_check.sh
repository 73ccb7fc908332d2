#!/bin/bash
# End-to-end check of the entry points on generated files (a few seconds each): image-GAN training from PNG files through
# the prefetcher's page-locked buffers, then the video inversion program on seeded random clips.
O=gpurun_out; mkdir -p $O
D=$(mktemp -d)
python - "$D" <<'PY'
import sys, os, numpy as np, cv2
d = os.path.join(sys.argv[1], "data", "faces"); os.makedirs(d)
rs = np.random.RandomState(0)
for i in range(40):
    cv2.imwrite(os.path.join(d, "f%03d.png" % i), rs.randint(0, 256, (80, 72, 3)).astype(np.uint8))
PY
t0=$(date +%s)
timeout ${CLI_TIMEOUT:-20} python gif-gan_b200/gifgan/main.py --dataset faces --data_dir $D/data --image_glob '*.png' --is_train true --epoch 1 \
  --batch_size 8 --image_size 64 --is_crop true --checkpoint_dir $D/ck --sample_dir $D/samples > $O/cli_main.log 2>&1
echo "main rc=$? t=$(( $(date +%s) - t0 ))"; grep -c "^Epoch" $O/cli_main.log; tail -n 2 $O/cli_main.log | cut -c1-200
timeout ${CLI_TIMEOUT:-20} python gif-gan_b200/gifgan/z_space_finder.py --synthetic 2 --video_batch_size 8 --vid_length 2 --num_initial_steps 6 \
  --num_steps_per_frame 3 --discriminator_mode inference --pixel_L2_weight 0.5 --output_z_folder $D/z --output_image_folder $D/img > $O/cli_zfinder.log 2>&1
echo "z_space_finder rc=$? t=$(( $(date +%s) - t0 ))"; tail -n 3 $O/cli_zfinder.log | cut -c1-200; ls $D/z $D/img 2>&1 | head
python - "$D" <<'PY'
import sys, os, numpy as np
z = np.load(os.path.join(sys.argv[1], "z", "synthetic_0000.npy")); print("latents", z.shape, float(np.abs(z).max()))
PY
rm -rf $D
