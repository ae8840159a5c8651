// qengine_cuda.hpp — drop-in replacement of the reference header of the same name
// (/root/reference/include/qengine_cuda.hpp) : class Qrack::QEngineCUDA on the B200-native state-vector core.
//
// The class keeps the reference's name, base class and shared positional constructor signature
// (reference include/qengine_cuda.hpp:280-284) so that include/qfactory.hpp:80,128,172,237, QPager, QHybrid, QUnit,
// QUnitMulti, the unit tests and the benchmarks compile and run UNCHANGED with ENABLE_CUDA=1.  It contains no CUDA
// code: every sweep over amplitudes is one call into the C ABI of include/b200sv.h (libb200sv.so, hand-written
// sm_100a kernels).  What stays here is QEngine-level bookkeeping only: runningNorm / doNormalize, argument checks that
// must throw std::invalid_argument, and the zero-state shortcuts.
#pragma once

#include "common/cudaengine.cuh"
#include "qengine.hpp"

struct b200sv_state;

namespace Qrack {

class QEngineCUDA;
typedef std::shared_ptr<QEngineCUDA> QEngineCUDAPtr;

class QEngineCUDA : public QEngine {
protected:
    b200sv_state* sv;
    int64_t deviceID;
    bool svCountSynced; // guards SetQubitCount() against re-creating the handle after a structural ABI call

    void Check(int rc) const;
    void ResizeZero(bitLenInt qb);
    void SyncQubitCount();
    QEngineCUDAPtr Cast(QInterfacePtr p, const char* what) const;

    /// QInterface::TryDecompose (src/qinterface/qinterface.cpp:836-853) adopts another instance's state through this
    /// protected hook; the reference engines share the buffer (include/qengine_cpu.hpp:47-55), here it is a device copy.
    using QEngine::Copy;
    void Copy(QInterfacePtr orig) { Copy(std::dynamic_pointer_cast<QEngineCUDA>(orig)); }
    void Copy(QEngineCUDAPtr orig);

    /// OR of 2^control; throws std::invalid_argument like ThrowIfQbIdArrayIsBad (common/qrack_functions.hpp)
    uint64_t CtrlMask(const std::vector<bitLenInt>& controls, const char* what) const;

public:
    /// 1 / OclMemDenom of device memory is the most a single state vector should take (test/benchmarks_main.cpp:288)
    static const bitCapIntOcl OclMemDenom = 3U;

    QEngineCUDA(bitLenInt qBitCount, const bitCapInt& initState, qrack_rand_gen_ptr rgp = nullptr,
        const complex& phaseFac = CMPLX_DEFAULT_ARG, bool doNorm = false, bool randomGlobalPhase = true,
        bool useHostMem = false, int64_t devID = -1, bool useHardwareRNG = true, bool ignored = false,
        real1_f norm_thresh = REAL1_EPSILON, std::vector<int64_t> ignored2 = {}, bitLenInt ignored4 = 0U,
        real1_f ignored3 = _qrack_qunit_sep_thresh);
    ~QEngineCUDA();

    bool isOpenCL() { return true; }
    void SetDevice(int64_t dID);
    int64_t GetDevice() { return deviceID; }
    bitCapIntOcl GetMaxSize();
    b200sv_state* Handle() { return sv; }

    void SetQubitCount(bitLenInt qb);
    void Finish();
    bool isFinished();
    void Dump() {}

    // ---- QEngine page / buffer virtuals (qengine.hpp:127-152) ----
    void ZeroAmplitudes();
    void CopyStateVec(QEnginePtr src);
    bool IsZeroAmplitude();
    void GetAmplitudePage(complex* pagePtr, bitCapIntOcl offset, bitCapIntOcl length);
    void SetAmplitudePage(const complex* pagePtr, bitCapIntOcl offset, bitCapIntOcl length);
    void SetAmplitudePage(QEnginePtr pageEnginePtr, bitCapIntOcl srcOffset, bitCapIntOcl dstOffset, bitCapIntOcl length);
    void ShuffleBuffers(QEnginePtr engine);
    QEnginePtr CloneEmpty();
    QInterfacePtr Clone();
    void QueueSetDoNormalize(bool doNorm) { doNormalize = doNorm; }
    void QueueSetRunningNorm(real1_f runningNrm) { runningNorm = (real1)runningNrm; }

    // ---- state access (qinterface.hpp:313-344) ----
    void SetQuantumState(const complex* inputState);
    void GetQuantumState(complex* outputState);
    void GetProbs(real1* outputProbs);
    complex GetAmplitude(const bitCapInt& perm);
    void SetAmplitude(const bitCapInt& perm, const complex& amp);
    void SetPermutation(const bitCapInt& perm, const complex& phaseFac = CMPLX_DEFAULT_ARG);

    // ---- the gate hot path ----
    void Apply2x2(bitCapInt offset1, bitCapInt offset2, const complex* mtrx, bitLenInt bitCount, bitCapInt const* qPowersSorted,
        bool doCalcNorm, real1_f norm_thresh = REAL1_DEFAULT_ARG);
    using QEngine::ApplyM;
    void ApplyM(const bitCapInt& regMask, const bitCapInt& result, const complex& nrm);
    void XMask(const bitCapInt& mask);
    void PhaseParity(real1_f radians, const bitCapInt& mask);
    void PhaseRootNMask(bitLenInt n, const bitCapInt& mask);
    using QEngine::UniformlyControlledSingleBit;
    void UniformlyControlledSingleBit(const std::vector<bitLenInt>& controls, bitLenInt qubitIndex, const complex* mtrxs,
        const std::vector<bitCapInt>& mtrxSkipPowers, const bitCapInt& mtrxSkipValueMask);
    void UniformParityRZ(const bitCapInt& mask, real1_f angle);
    void CUniformParityRZ(const std::vector<bitLenInt>& controls, const bitCapInt& mask, real1_f angle);

    // ---- probabilities / measurement ----
    real1_f Prob(bitLenInt qubit);
    real1_f CtrlOrAntiProb(bool controlState, bitLenInt control, bitLenInt target);
    real1_f ProbReg(bitLenInt start, bitLenInt length, const bitCapInt& permutation);
    real1_f ProbMask(const bitCapInt& mask, const bitCapInt& permutation);
    void ProbMaskAll(const bitCapInt& mask, real1* probsArray);
    void ProbRegAll(bitLenInt start, bitLenInt length, real1* probsArray);
    real1_f ProbParity(const bitCapInt& mask);
    bool ForceMParity(const bitCapInt& mask, bool result, bool doForce = true);
    bitCapInt MAll();
    using QInterface::HighestProbAll;
    bitCapInt HighestProbAll(); // device arg-max; the QInterface default asks ProbAll() for every permutation
    real1_f FirstNonzeroPhase() { return IsZeroAmplitude() ? ZERO_R1_F : QInterface::FirstNonzeroPhase(); }
    real1_f GetExpectation(bitLenInt valueStart, bitLenInt valueLength);

    // ---- structure ----
    using QEngine::Compose;
    bitLenInt Compose(QEngineCUDAPtr toCopy);
    bitLenInt Compose(QInterfacePtr toCopy) { return Compose(Cast(toCopy, "Compose")); }
    bitLenInt Compose(QEngineCUDAPtr toCopy, bitLenInt start);
    bitLenInt Compose(QInterfacePtr toCopy, bitLenInt start) { return Compose(Cast(toCopy, "Compose"), start); }
    using QEngine::Decompose;
    void Decompose(bitLenInt start, QInterfacePtr dest);
    void Dispose(bitLenInt start, bitLenInt length);
    void Dispose(bitLenInt start, bitLenInt length, const bitCapInt& disposedPerm);
    using QEngine::Allocate;
    bitLenInt Allocate(bitLenInt start, bitLenInt length);

    // ---- norm ----
    real1_f SumSqrDiff(QInterfacePtr toCompare);
    void NormalizeState(
        real1_f nrm = REAL1_DEFAULT_ARG, real1_f norm_thresh = REAL1_DEFAULT_ARG, real1_f phaseArg = ZERO_R1_F);
    void UpdateRunningNorm(real1_f norm_thresh = REAL1_DEFAULT_ARG);

    void ROL(bitLenInt shift, bitLenInt start, bitLenInt length);
    void ROR(bitLenInt shift, bitLenInt start, bitLenInt length);

#if ENABLE_ALU
    // ---- QAlu (include/qalu.hpp): one device sweep each (include/b200sv.h "QAlu family") ----
    void INC(const bitCapInt& toAdd, bitLenInt start, bitLenInt length);
    void CINC(const bitCapInt& toAdd, bitLenInt start, bitLenInt length, const std::vector<bitLenInt>& controls);
    void INCDECC(const bitCapInt& toMod, bitLenInt start, bitLenInt length, bitLenInt carryIndex);
    void INCS(const bitCapInt& toAdd, bitLenInt start, bitLenInt length, bitLenInt overflowIndex);
    void MULModNOut(const bitCapInt& toMod, const bitCapInt& modN, bitLenInt inStart, bitLenInt outStart, bitLenInt length);
    void IMULModNOut(const bitCapInt& toMod, const bitCapInt& modN, bitLenInt inStart, bitLenInt outStart, bitLenInt length);
    void CMULModNOut(const bitCapInt& toMod, const bitCapInt& modN, bitLenInt inStart, bitLenInt outStart, bitLenInt length,
        const std::vector<bitLenInt>& controls);
    void CIMULModNOut(const bitCapInt& toMod, const bitCapInt& modN, bitLenInt inStart, bitLenInt outStart, bitLenInt length,
        const std::vector<bitLenInt>& controls);
    void PhaseFlipIfLess(const bitCapInt& greaterPerm, bitLenInt start, bitLenInt length);
    void CPhaseFlipIfLess(const bitCapInt& greaterPerm, bitLenInt start, bitLenInt length, bitLenInt flagIndex);
    void INCDECSC(const bitCapInt& toMod, bitLenInt start, bitLenInt length, bitLenInt carryIndex);
    void INCDECSC(const bitCapInt& toMod, bitLenInt start, bitLenInt length, bitLenInt overflowIndex, bitLenInt carryIndex);
    void MUL(const bitCapInt& toMul, bitLenInt start, bitLenInt carryStart, bitLenInt length);
    void DIV(const bitCapInt& toDiv, bitLenInt start, bitLenInt carryStart, bitLenInt length);
    void POWModNOut(const bitCapInt& base, const bitCapInt& modN, bitLenInt inStart, bitLenInt outStart, bitLenInt length);
    void CMUL(const bitCapInt& toMul, bitLenInt start, bitLenInt carryStart, bitLenInt length,
        const std::vector<bitLenInt>& controls);
    void CDIV(const bitCapInt& toDiv, bitLenInt start, bitLenInt carryStart, bitLenInt length,
        const std::vector<bitLenInt>& controls);
    void CPOWModNOut(const bitCapInt& base, const bitCapInt& modN, bitLenInt inStart, bitLenInt outStart, bitLenInt length,
        const std::vector<bitLenInt>& controls);
    bitCapInt IndexedLDA(bitLenInt indexStart, bitLenInt indexLength, bitLenInt valueStart, bitLenInt valueLength,
        const unsigned char* values, bool resetValue = true);
    bitCapInt IndexedADC(bitLenInt indexStart, bitLenInt indexLength, bitLenInt valueStart, bitLenInt valueLength,
        bitLenInt carryIndex, const unsigned char* values);
    bitCapInt IndexedSBC(bitLenInt indexStart, bitLenInt indexLength, bitLenInt valueStart, bitLenInt valueLength,
        bitLenInt carryIndex, const unsigned char* values);
    void Hash(bitLenInt start, bitLenInt length, const unsigned char* values);
#endif
};

} // namespace Qrack
