N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tests/_ddp_gpu_worker.py 2>&1 | grep -E "DDP_|Error|error|assert|Traceback" | sort | uniq -c | tail -12
for v in "ABCGPT_DDP_NVLS=1" "ABCGPT_DDP_NVLS=0" "ABCGPT_DDP_NVLS=1"; do env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 6 --no-cpu-baseline > gpurun_out/s7_n${N}.json 2> gpurun_out/s7_n${N}.err; python -c "
import json
for l in open('gpurun_out/s7_n${N}.json'):
    if l.startswith('{'):
        d=json.loads(l); print('cfg3 n$N $v',round(d['ms_per_step'],3),round(d['value']),d['loss_after'],d['config']['grad_exchange'][:6])"; grep -E "Error|Traceback|Warning" gpurun_out/s7_n${N}.err | head -3; cp gpurun_out/s7_n${N}.json gpurun_out/s7_n${N}_$(echo $v | tr -d 'A-Z_= ').json; done
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g python bench.py --steps 20 --warmup 6 --no-cpu-baseline > gpurun_out/s7_g$g.json 2>/dev/null & done
wait
python -c "
import json
print('independent', [round(json.load(open(f'gpurun_out/s7_g{g}.json'))['ms_per_step'],3) for g in range($N)])"
