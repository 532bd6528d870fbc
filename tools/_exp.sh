cd $GRAFT_REPO_ROOT
python tools/time_features.py 2>&1 | tail -1
python tools/calibrate_kappa.py 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu -k "features or engines or golden or f64" 2>&1 | tail -3
