# usage: source profiles/gpu_health.sh; health "label"  -- prints whether the GPU still answers
health() {
  if timeout 60 python -c "import torch; x=torch.ones(1<<20,device='cuda'); torch.cuda.synchronize(); print('GPU ok after $1:', float(x.sum()))" 2>&1 | tail -1; then :; else echo "GPU NOT ANSWERING after $1"; fi
  timeout 20 nvidia-smi --query-gpu=name,clocks.sm,temperature.gpu --format=csv,noheader 2>&1 | tail -1
}
