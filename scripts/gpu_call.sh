echo "normal"; timeout 200 python scripts/time_roles.py | head -1
echo "opponent kernel without its fc2 stream (results invalid)"; COEVONET_LIB=$PWD/scripts/probe/exp2_libcoevonet_b200.so timeout 200 python scripts/time_roles.py | head -1
echo "member only"; CEV_LS_SKIP=1 timeout 200 python scripts/time_roles.py | head -1
echo "opp only"; CEV_LS_SKIP=2 timeout 200 python scripts/time_roles.py | head -1
