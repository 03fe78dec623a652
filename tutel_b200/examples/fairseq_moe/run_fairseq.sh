#!/bin/bash -e
# fairseq language-model training with tutel_b200 MoE layers on all GPUs of this node.
#
#   MOE=2 L_AUX_WT=0.01 FP16=1 ./run_fairseq.sh /path/to/data-bin/wikitext-103
#
# MOE=<n>        every n-th transformer layer gets an MoE feed-forward (integration.convert_transformer_layers)
# L_AUX_WT=<w>   weight of the load-balancing loss;   NO_OVERFLOW=1 zeroes inf gradients;   FP16=1 mixed precision
# NGPU=<n>       number of local GPUs (default: all);  FLAGS="..." replaces the precision flags
# The recipe (task, architecture, optimiser schedule) is the one of the reference's tutel/examples/fairseq_moe/run_fairseq.sh;
# fairseq must be importable - train_moe.py next to this script wraps fairseq_cli.train.
here=$(cd "$(dirname "$0")" && pwd)
data=${1:-./wikitext-103}
[ $# -gt 0 ] && shift
ngpu=${NGPU:-$(nvidia-smi -L | wc -l)}

precision=()
if [[ -n "$FLAGS" ]]; then
    read -r -a precision <<< "$FLAGS"
elif [[ "$FP16" == "1" ]]; then
    precision=(--fp16 --fp16-init-scale 4 --fp16-no-flatten-grads)
fi

model=(--task language_modeling --arch transformer_lm_gpt2_tiny --tokens-per-sample 256 --batch-size 8)
optim=(--optimizer adam --adam-betas "(0.9,0.98)" --lr 0.0001 --lr-scheduler inverse_sqrt --warmup-updates 4000 --max-update 500000)
runtime=(--ddp-backend legacy_ddp --log-format json --log-interval 100 --save-dir ./fairseq_checkpoints)

exec python3 -m torch.distributed.run --nnodes=1 --nproc-per-node="$ngpu" --master-addr 127.0.0.1 \
    "$here/train_moe.py" "$data" "${model[@]}" "${optim[@]}" "${runtime[@]}" "${precision[@]}" "$@"
