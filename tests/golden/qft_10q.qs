qubits 10
SetPermutation 241
H 0
H 2
H 3
H 4
H 6
H 8
QFT 0 10
T 3
IQFT 1 8
Prob 0
Prob 1
Prob 2
Prob 3
Prob 4
Prob 5
Prob 6
Prob 7
Prob 8
Prob 9
