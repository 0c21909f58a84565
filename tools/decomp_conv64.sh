for res in "" 1; do for dbg in 0 1 2 4; do echo "res=$res dbg=$dbg"; PROF_RES=$res PROF_STATS=1 WSR_TC_DBG=$dbg python tools/prof_conv.py 64 64 64 128 256 3 10; done; done
echo "unstaged:"; for res in "" 1; do PROF_RES=$res PROF_STATS=1 WSR_NO_STAGE=1 python tools/prof_conv.py 64 64 64 128 256 3 10; done
PROF_STATS=1 python tools/prof_conv.py 64 192 64 128 256 3 10
PROF_STATS=1 python tools/prof_conv.py 64 128 128 64 128 3 10
PROF_STATS=1 WSR_TC_DBG=2 python tools/prof_conv.py 64 128 128 64 128 3 10
PROF_STATS=1 WSR_TC_DBG=1 python tools/prof_conv.py 64 128 128 64 128 3 10
