export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR --nproc-per-node 8 tests/_mgpu_worker.py 2>gpurun_out/w8.err | grep sharded
run() { n=$1; shift; name=$1; shift; timeout 200 $TR --nproc-per-node $n tools/bench_sharded.py --steps 10 "$@" 2>gpurun_out/sh.err > gpurun_out/bench_sharded_$name.json || tail -5 gpurun_out/sh.err; python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(d['n_gpus'], d['exchange'], round(d['ms_per_step'],3)); [print(p['device_ms']) for p in d.get('phases_per_rank',[])[:2]]" gpurun_out/bench_sharded_$name.json; }
run 8 n8_gather
run 8 n8_peer --peer
run 8 n8_peer_phases --peer --phases
run 4 n4_peer --peer
