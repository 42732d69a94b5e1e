run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); n=d['newton']; print('solve %.4f s its %d  ms/it %.4f launches %d asm %.2f'%(n['solve_s'],n['krylov_iterations'],n['krylov_ms']/n['krylov_iterations'],n['gpu_launches'],n['assemble_ms']))"; }
run BP_X=1
run BP_P2P=0
run BP_MG_REPL=0
run BP_MG_REPL=200000
run BP_MG_COUPLED_GRAPH=0
run BP_MG_COUPLED=0
