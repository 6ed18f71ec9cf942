# A/B of data-parallel variants on one box: tools/ddp_ab.sh N "ENV=.. ENV=.." "ENV=.." ...   (each argument after N = one run's environment)
N=$1; shift
i=0
for v in "$@"; do i=$((i+1)); env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/ddp_ab_$i.json 2> gpurun_out/ddp_ab_$i.err; python -c "
import json
for l in open('gpurun_out/ddp_ab_$i.json'):
    if l.startswith('{'):
        d=json.loads(l); kb=d['kernel_breakdown']; print('n$N [$v]',round(d['ms_per_step'],3),round(d['value']),d['loss_after'],d['config']['grad_exchange'][:4],{k:x['ms'] for k,x in kb.items() if k in('nvls_allreduce_sumsq','gemm_dgrad','gemm_wgrad','attn_bwd')})"; grep -E "Error|Traceback" gpurun_out/ddp_ab_$i.err | head -2; done
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g python bench.py --steps 20 --warmup 6 --no-cpu-baseline > gpurun_out/ddp_ab_g$g.json 2>/dev/null & done
wait
python -c "
import json
print('independent', [round(json.load(open(f'gpurun_out/ddp_ab_g{g}.json'))['ms_per_step'],3) for g in range($N)])"
