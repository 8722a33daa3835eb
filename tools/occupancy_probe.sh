for d in 0 38912 97280; do
  SG_HOT_DYNSMEM=$d SPACEGYM_B200_LIB=$PWD/spacegym_b200/libsg_dev_base.so python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-f32 --e2e-seconds 0.1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dyn $d', 'value %.4g'%d['value'], 'single %.4g'%d['single_handle']['value'], d['kernel_times']['hot_us'])"
done
