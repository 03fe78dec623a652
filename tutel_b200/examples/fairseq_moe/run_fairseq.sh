#!/bin/bash -e
# Language-model training with fairseq + tutel_b200 MoE layers on the GPUs of one node
# (same task / architecture as the reference's tutel/examples/fairseq_moe/run_fairseq.sh).
#   MOE=2 L_AUX_WT=0.01 FP16=1 ./run_fairseq.sh /path/to/data-bin/wikitext-103
# fairseq must be importable; `train_moe.py` next to this script wraps fairseq_cli.train (see README.md).
HERE=$(cd "$(dirname "$0")" && pwd)
NGPU=${NGPU:-$(nvidia-smi -L | wc -l)}
if [[ "$FP16" == "1" ]]; then
    FLAGS=${FLAGS:---fp16 --fp16-init-scale 4 --fp16-no-flatten-grads}
fi
python3 -m torch.distributed.run --nnodes=1 --nproc-per-node="$NGPU" --master-addr 127.0.0.1 \
    "$HERE/train_moe.py" "${@:-./wikitext-103}" \
    --ddp-backend legacy_ddp \
    --task language_modeling --tokens-per-sample 256 --batch-size 8 \
    --arch transformer_lm_gpt2_tiny \
    --optimizer adam --adam-betas "(0.9,0.98)" \
    --lr 0.0001 --lr-scheduler inverse_sqrt --warmup-updates 4000 \
    --max-update 500000 --log-format json --log-interval 100 \
    ${FLAGS} \
    --save-dir ./fairseq_checkpoints
