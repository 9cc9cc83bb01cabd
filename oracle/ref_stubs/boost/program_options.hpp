// empty: the reference CLI (main.cpp) is not compiled by the oracle build
