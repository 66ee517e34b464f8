#!/bin/bash
# BASELINE configs[3]: N_FRAMES synthetic 512x512 radar PNGs with one shared style image through `python -m ist_b200.main`
# on NGPU GPUs of this box (torchrun for NGPU > 1): decode / upload / optimise / download / encode pipelined per GPU, frames
# sharded round-robin, one final gather. Prints the summary line (frames/s incl. model load, style target, I/O, gather).
# usage: bash tools/run_frames_job.sh NGPU [N_FRAMES=256] [FRAMES_PER_BATCH=4] [TAG]
NGPU=${1:-1}; NF=${2:-256}; FPB=${3:-4}; TAG=${4:-r02}
D=/tmp/ist_frames_job
[ -f $D/vgg_conv.pth ] || python tools/make_frames.py $D $NF > /dev/null
OUT=$D/out_${NGPU}
rm -rf $OUT
ARGS="--content-dir $D/radar --style-img $D/style.png --output-dir $OUT --frames-per-batch $FPB --max-frames $NF --summary-json gpurun_out/${TAG}_frames_${NGPU}gpu.json MODEL.WEIGHTS $D/vgg_conv.pth"
if [ "$NGPU" = "1" ]; then
    python -m ist_b200.main $ARGS > gpurun_out/${TAG}_frames_${NGPU}gpu.log 2>&1
else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29533 -m ist_b200.main $ARGS > gpurun_out/${TAG}_frames_${NGPU}gpu.log 2>&1
fi
echo "rc=$? files=$(ls $OUT/*.png 2>/dev/null | wc -l) sha_of_files=$(cat $OUT/*.png | sha256sum | cut -c1-16)"
cat gpurun_out/${TAG}_frames_${NGPU}gpu.json
