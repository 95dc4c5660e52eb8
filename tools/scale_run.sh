set -x
for N in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 1000 --warmup 5 2>gpurun_out/b$N.err | tail -1 > gpurun_out/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/sc_bench.py --gpus $N --queries 500 2>gpurun_out/sc$N.err | tail -1 > gpurun_out/sc_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/sequence_bench.py --gpus $N --frames 200 2>gpurun_out/seq$N.err | tail -1 > gpurun_out/seq_n$N.json
done
for N in 4 2; do
CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((N-1))) python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $N --steps 1000 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_n$N.json
CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((N-1))) python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29525 tools/sc_bench.py --gpus $N --queries 500 2>/dev/null | tail -1 > gpurun_out/sc_n$N.json
CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((N-1))) python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29526 tools/sequence_bench.py --gpus $N --frames 200 2>/dev/null | tail -1 > gpurun_out/seq_n$N.json
done
python bench.py --steps 1000 --no-cpu --no-sweep 2>/dev/null | tail -1 > gpurun_out/bench_n1.json
python tools/sc_bench.py --queries 500 2>/dev/null | tail -1 > gpurun_out/sc_n1.json
python tools/sequence_bench.py --frames 200 --oracle-frames 0 2>/dev/null | tail -1 > gpurun_out/seq_n1.json
for f in gpurun_out/bench_n*.json gpurun_out/sc_n*.json gpurun_out/seq_n*.json; do echo $f; cut -c1-160 $f; done
