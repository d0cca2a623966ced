# One gpurun call that produces the round's evidence: tests, the bench line, the ncu launch list of the same bench command (eager,
# 2 diffusion steps), a full ncu capture of the dominant kernel and one of the IGSO3 sampler.  usage: bash scripts/capture_round_profiles.sh <tag>
TAG=${1:-r4}
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
SE3DIFF_B200_CUDA_GRAPH=0 SE3DIFF_B200_MODEL_GRAPH=0 python bench.py --diffusion-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-extras > /dev/null 2>&1 && \
SE3DIFF_B200_CUDA_GRAPH=0 SE3DIFF_B200_MODEL_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --diffusion-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-extras > gpurun_out/${TAG}_ncu_list.log 2>&1
python scripts/run_ipa_tc_once.py && ncu --set full --clock-control none --import-source on -k regex:k_ipa_tc -c 4 -o gpurun_out/${TAG}_ipa_tc python scripts/run_ipa_tc_once.py > gpurun_out/${TAG}_ncu_full.log 2>&1
python scripts/run_sampler_once.py && ncu --set full --clock-control none --import-source on -k regex:k_sample -c 4 -o gpurun_out/${TAG}_sampler python scripts/run_sampler_once.py > gpurun_out/${TAG}_ncu_sampler.log 2>&1
ls -la gpurun_out | tail -8
