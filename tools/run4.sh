set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PT="python -m pytest -q --timeout=150 --timeout-method=thread"
timeout 200 $PT tests/test_fusion_gpu.py -m gpu 2>&1 | tail -8 > gpurun_out/r02_pt4_fusion.log; cat gpurun_out/r02_pt4_fusion.log
timeout 700 $PT tests -m gpu --deselect tests/test_fusion_gpu.py 2>&1 | tail -30 > gpurun_out/r02_pt4_all.log; cat gpurun_out/r02_pt4_all.log
for v in default ring chainv3; do
  case $v in default) E="";; ring) E="SGB_FWD_RING=1";; chainv3) E="SGB_CHAIN_V3=1";; esac
  env $E timeout 150 python bench.py --steps 16 --warmup 3 --no-baselines --quick > gpurun_out/r02_b4_k3_$v.json 2> gpurun_out/r02_b4_k3_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/r02_b4_k3_$v.json'));print('$v',d['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
done
SGB_BLEND_MMA=1 timeout 120 $PT tests/test_parity_gpu.py -k "channel_forward_vs_reference" 2>&1 | tail -15 > gpurun_out/r02_pt4_mma.log; cat gpurun_out/r02_pt4_mma.log
SGB_BLEND_MMA=1 timeout 150 python bench.py --steps 16 --warmup 3 --no-baselines --quick > gpurun_out/r02_b4_k3_mma.json 2> gpurun_out/r02_b4_k3_mma.err
python -c "
import json;d=json.load(open('gpurun_out/r02_b4_k3_mma.json'));print('mma',d['ms_per_step'],{k:round(x,3) for k,x in d['stage_ms'].items()})"
SGB_K5_VIEWS=24 timeout 200 python bench.py --config K5 --steps 2 --no-baselines > gpurun_out/r02_b4_k5_v24.json 2> gpurun_out/r02_b4_k5_v24.err
tail -c 800 gpurun_out/r02_b4_k5_v24.err
python -c "
import json;d=json.load(open('gpurun_out/r02_b4_k5_v24.json'));print('k5',d['ms_per_step'],d['stage_ms'],d['config']['N_vis_per_view'],d['roofline'])"
ls -la gpurun_out | tail -14
