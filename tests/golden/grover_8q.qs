qubits 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
DEC 3 0 8
ZeroPhaseFlip 0 8
INC 3 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
ZeroPhaseFlip 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
DEC 3 0 8
ZeroPhaseFlip 0 8
INC 3 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
ZeroPhaseFlip 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
DEC 3 0 8
ZeroPhaseFlip 0 8
INC 3 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
ZeroPhaseFlip 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
DEC 3 0 8
ZeroPhaseFlip 0 8
INC 3 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
ZeroPhaseFlip 0 8
H 0
H 1
H 2
H 3
H 4
H 5
H 6
H 7
ProbAll 3
