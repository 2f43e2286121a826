echo "normal, opp only, 148 SMs"; CEV_LS_FORK=0 CEV_LS_SKIP=2 timeout 120 python scripts/time_ls.py
echo "producers write nothing (results invalid)"; COEVONET_LIB=$PWD/scripts/probe/exp1_libcoevonet_b200.so CEV_LS_FORK=0 CEV_LS_SKIP=2 timeout 120 python scripts/time_ls.py
