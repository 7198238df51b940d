for f in 1 2 4; do for b in 1 2 4; do
GSR_FWD_PPT=$f GSR_BWD_PPT=$b timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ppt_${f}_${b}.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/ppt_${f}_${b}.json')); s=d['stages']; print('fwdPPT=$f bwdPPT=$b step', round(d['ms_per_step'],3), 'render_fwd', round(s['render_fwd']['ms'],3), 'render_bwd', round(s['render_bwd']['ms'],3), 'pre', round(s['preprocess_fwd']['ms'],3))"
done; done
