// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Replays a "qscript" (the circuit-script text format documented in
// qrack_b200/qscript.py) on the UNMODIFIED reference engine through the
// reference's own public API (Qrack::CreateQuantumInterface -> QEngineCPU, or
// QPager-over-QEngineCPU), dumps the final state vector(s) and the results of
// the query ops, and optionally times the replay.  It is linked against
// oracle/_ref/f{32,64}/libqrack.a, which oracle/Makefile compiles straight
// from /root/reference/src with g++.
//
// Reference API used: include/qfactory.hpp:48-258 (CreateQuantumInterface),
// include/qinterface.hpp (gate sugar), include/qengine.hpp (QEngine).
//
//   ref_harness <script> [--dump PREFIX] [--results FILE] [--engine cpu|pager:<qpp>]
//               [--threads N] [--time]
#include "qfactory.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace Qrack;

static std::string g_engine = "cpu";
static int g_threads = 0;

static QInterfacePtr make_reg(bitLenInt n, uint64_t perm)
{
    qrack_rand_gen_ptr rng = std::make_shared<qrack_rand_gen>();
    rng->seed(20250921U);
    QInterfacePtr q;
    if (g_engine.rfind("pager:", 0) == 0) {
        const int qpp = atoi(g_engine.c_str() + 6);
        q = CreateQuantumInterface({ QINTERFACE_QPAGER, QINTERFACE_CPU }, n, ZERO_BCI, rng, ONE_CMPLX, false, false,
            false, -1, false, false, REAL1_EPSILON, std::vector<int64_t>{}, (bitLenInt)qpp);
#if ENABLE_CUDA
    } else if (g_engine == "cuda") {
        // the QEngineCUDA slot (dropin/): same factory call, different enum
        q = CreateQuantumInterface(QINTERFACE_CUDA, n, ZERO_BCI, rng, ONE_CMPLX, false, false, false, -1, false);
    } else if (g_engine.rfind("pager-cuda:", 0) == 0) {
        const int qpp = atoi(g_engine.c_str() + 11);
        q = CreateQuantumInterface({ QINTERFACE_QPAGER, QINTERFACE_CUDA }, n, ZERO_BCI, rng, ONE_CMPLX, false, false,
            false, -1, false, false, REAL1_EPSILON, std::vector<int64_t>{}, (bitLenInt)qpp);
    } else if (g_engine == "hybrid") {
        q = CreateQuantumInterface(QINTERFACE_HYBRID, n, ZERO_BCI, rng, ONE_CMPLX, false, false, false, -1, false);
    } else if (g_engine == "qunit-cuda") {
        q = CreateQuantumInterface({ QINTERFACE_QUNIT, QINTERFACE_CUDA }, n, ZERO_BCI, rng, ONE_CMPLX, false, false, false, -1, false);
#endif
    } else {
        q = CreateQuantumInterface(QINTERFACE_CPU, n, ZERO_BCI, rng, ONE_CMPLX, false, false, false, -1, false);
    }
    // QPager ignores initState at this commit (src/qpager.cpp:54): always set explicitly.
    q->SetPermutation(bitCapInt(perm), ONE_CMPLX);
    if (g_threads > 0) {
        q->SetConcurrency((unsigned)g_threads);
    }
    return q;
}

struct Tok {
    std::vector<std::string> t;
    size_t p = 0;
    std::string s() { return t.at(p++); }
    long long i() { return std::stoll(t.at(p++)); }
    uint64_t u() { return std::stoull(t.at(p++)); }
    double d() { return std::stod(t.at(p++)); }
    complex c()
    {
        const double re = d();
        const double im = d();
        return complex((real1)re, (real1)im);
    }
    std::vector<bitLenInt> qubits()
    {
        const long long n = i();
        std::vector<bitLenInt> v;
        for (long long k = 0; k < n; ++k) {
            v.push_back((bitLenInt)i());
        }
        return v;
    }
    std::vector<unsigned char> hex() // classical table as one hex token
    {
        const std::string h = s();
        std::vector<unsigned char> v;
        for (size_t k = 0; k + 1 < h.size(); k += 2) {
            v.push_back((unsigned char)std::stoul(h.substr(k, 2), nullptr, 16));
        }
        return v;
    }
    void mtrx(complex* m)
    {
        for (int k = 0; k < 4; ++k) {
            m[k] = c();
        }
    }
};

int main(int argc, char** argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <script> [--dump PREFIX] [--results FILE] [--engine cpu|pager:<qpp>] [--threads N] [--time]\n", argv[0]);
        return 2;
    }
    std::string script = argv[1], dumpPrefix, resultsFile;
    bool doTime = false;
    for (int a = 2; a < argc; ++a) {
        std::string s = argv[a];
        if (s == "--dump" && a + 1 < argc) {
            dumpPrefix = argv[++a];
        } else if (s == "--results" && a + 1 < argc) {
            resultsFile = argv[++a];
        } else if (s == "--engine" && a + 1 < argc) {
            g_engine = argv[++a];
        } else if (s == "--threads" && a + 1 < argc) {
            g_threads = atoi(argv[++a]);
        } else if (s == "--time") {
            doTime = true;
        }
    }

    std::ifstream in(script);
    if (!in.good()) {
        fprintf(stderr, "cannot open %s\n", script.c_str());
        return 2;
    }
    std::vector<Tok> lines;
    std::string line;
    while (std::getline(in, line)) {
        const size_t h = line.find('#');
        if (h != std::string::npos) {
            line = line.substr(0, h);
        }
        std::istringstream ss(line);
        Tok tk;
        std::string w;
        while (ss >> w) {
            tk.t.push_back(w);
        }
        if (!tk.t.empty()) {
            lines.push_back(tk);
        }
    }

    std::map<int, QInterfacePtr> regs;
    std::vector<std::string> results;
    char buf[256];
    size_t gateCount = 0, totalGates = 0;
    double seconds = 0;
    std::vector<std::pair<size_t, double>> segments; // (gates, seconds) of every TIC..TOC region
    auto t0 = std::chrono::high_resolution_clock::now();
    bool timing = false;

    for (Tok& tk : lines) {
        int r = 0;
        if (tk.t[0][0] == '@') {
            r = atoi(tk.t[0].c_str() + 1);
            tk.p = 1;
        }
        const std::string op = tk.s();
        if (op == "qubits") {
            regs[0] = make_reg((bitLenInt)tk.i(), 0U);
            continue;
        }
        if (op == "reg") {
            const int id = (int)tk.i();
            const bitLenInt n = (bitLenInt)tk.i();
            const uint64_t perm = tk.u();
            regs[id] = make_reg(n, perm);
            continue;
        }
        if (op == "TIC") {
            for (auto& kv : regs) {
                kv.second->Finish();
            }
            timing = true;
            gateCount = 0;
            t0 = std::chrono::high_resolution_clock::now();
            continue;
        }
        if (op == "TOC") {
            for (auto& kv : regs) {
                kv.second->Finish();
            }
            const double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
            seconds += dt;
            totalGates += gateCount;
            segments.push_back(std::make_pair(gateCount, dt));
            timing = false;
            continue;
        }
        QInterfacePtr q = regs.at(r);
        ++gateCount;
        complex m[4];
        if (op == "H") {
            q->H((bitLenInt)tk.i());
        } else if (op == "X") {
            q->X((bitLenInt)tk.i());
        } else if (op == "Y") {
            q->Y((bitLenInt)tk.i());
        } else if (op == "Z") {
            q->Z((bitLenInt)tk.i());
        } else if (op == "S") {
            q->S((bitLenInt)tk.i());
        } else if (op == "IS") {
            q->IS((bitLenInt)tk.i());
        } else if (op == "T") {
            q->T((bitLenInt)tk.i());
        } else if (op == "IT") {
            q->IT((bitLenInt)tk.i());
        } else if (op == "SqrtX") {
            q->SqrtX((bitLenInt)tk.i());
        } else if (op == "CNOT") {
            const bitLenInt c = tk.i(), t = tk.i();
            q->CNOT(c, t);
        } else if (op == "AntiCNOT") {
            const bitLenInt c = tk.i(), t = tk.i();
            q->AntiCNOT(c, t);
        } else if (op == "CZ") {
            const bitLenInt c = tk.i(), t = tk.i();
            q->CZ(c, t);
        } else if (op == "CY") {
            const bitLenInt c = tk.i(), t = tk.i();
            q->CY(c, t);
        } else if (op == "CCNOT") {
            const bitLenInt c1 = tk.i(), c2 = tk.i(), t = tk.i();
            q->CCNOT(c1, c2, t);
        } else if (op == "Swap") {
            const bitLenInt a = tk.i(), b = tk.i();
            q->Swap(a, b);
        } else if (op == "ISwap") {
            const bitLenInt a = tk.i(), b = tk.i();
            q->ISwap(a, b);
        } else if (op == "SqrtSwap") {
            const bitLenInt a = tk.i(), b = tk.i();
            q->SqrtSwap(a, b);
        } else if (op == "FSim") {
            const double th = tk.d(), ph = tk.d();
            const bitLenInt a = tk.i(), b = tk.i();
            q->FSim((real1_f)th, (real1_f)ph, a, b);
        } else if (op == "CSwap") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt a = tk.i(), b = tk.i();
            q->CSwap(c, a, b);
        } else if (op == "AntiCSwap") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt a = tk.i(), b = tk.i();
            q->AntiCSwap(c, a, b);
        } else if (op == "U") {
            const bitLenInt t = tk.i();
            const double th = tk.d(), ph = tk.d(), la = tk.d();
            q->U(t, (real1_f)th, (real1_f)ph, (real1_f)la);
        } else if (op == "AI") {
            const bitLenInt t = tk.i();
            const double az = tk.d(), inc = tk.d();
            q->AI(t, (real1_f)az, (real1_f)inc);
        } else if (op == "IAI") {
            const bitLenInt t = tk.i();
            const double az = tk.d(), inc = tk.d();
            q->IAI(t, (real1_f)az, (real1_f)inc);
        } else if (op == "Phase") {
            const bitLenInt t = tk.i();
            const complex a = tk.c(), b = tk.c();
            q->Phase(a, b, t);
        } else if (op == "Invert") {
            const bitLenInt t = tk.i();
            const complex a = tk.c(), b = tk.c();
            q->Invert(a, b, t);
        } else if (op == "Mtrx") {
            const bitLenInt t = tk.i();
            tk.mtrx(m);
            q->Mtrx(m, t);
        } else if (op == "MCMtrx") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            tk.mtrx(m);
            q->MCMtrx(c, m, t);
        } else if (op == "MACMtrx") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            tk.mtrx(m);
            q->MACMtrx(c, m, t);
        } else if (op == "UCMtrx") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            const uint64_t perm = tk.u();
            tk.mtrx(m);
            q->UCMtrx(c, m, t, bitCapInt(perm));
        } else if (op == "MCPhase") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            const complex a = tk.c(), b = tk.c();
            q->MCPhase(c, a, b, t);
        } else if (op == "MCInvert") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            const complex a = tk.c(), b = tk.c();
            q->MCInvert(c, a, b, t);
        } else if (op == "MACPhase") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            const complex a = tk.c(), b = tk.c();
            q->MACPhase(c, a, b, t);
        } else if (op == "MACInvert") {
            std::vector<bitLenInt> c = tk.qubits();
            const bitLenInt t = tk.i();
            const complex a = tk.c(), b = tk.c();
            q->MACInvert(c, a, b, t);
        } else if (op == "PhaseRootN") {
            const bitLenInt n = tk.i(), t = tk.i();
            q->PhaseRootN(n, t);
        } else if (op == "CPhaseRootN") {
            const bitLenInt n = tk.i(), c = tk.i(), t = tk.i();
            q->CPhaseRootN(n, c, t);
        } else if (op == "QFT") {
            const bitLenInt s = tk.i(), l = tk.i();
            q->QFT(s, l);
        } else if (op == "IQFT") {
            const bitLenInt s = tk.i(), l = tk.i();
            q->IQFT(s, l);
        } else if (op == "XMask") {
            q->XMask(bitCapInt(tk.u()));
        } else if (op == "ZMask") {
            q->ZMask(bitCapInt(tk.u()));
        } else if (op == "PhaseParity") {
            const double rad = tk.d();
            q->PhaseParity((real1_f)rad, bitCapInt(tk.u()));
        } else if (op == "PhaseRootNMask") {
            const bitLenInt n = tk.i();
            q->PhaseRootNMask(n, bitCapInt(tk.u()));
        } else if (op == "ZeroPhaseFlip") {
            const bitLenInt s = tk.i(), l = tk.i();
            q->ZeroPhaseFlip(s, l);
        } else if (op == "INC") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i();
            q->INC(bitCapInt(v), s, l);
        } else if (op == "DEC") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i();
            q->DEC(bitCapInt(v), s, l);
        } else if (op == "ROL" || op == "ROR") {
            const bitLenInt sh = tk.i(), s = tk.i(), l = tk.i();
            if (op == "ROL") {
                q->ROL(sh, s, l);
            } else {
                q->ROR(sh, s, l);
            }
#if ENABLE_ALU
        } else if (op == "CINC" || op == "CDEC") {
            std::vector<bitLenInt> c = tk.qubits();
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i();
            if (op == "CINC") {
                q->CINC(bitCapInt(v), s, l, c);
            } else {
                q->CDEC(bitCapInt(v), s, l, c);
            }
        } else if (op == "INCC" || op == "DECC" || op == "INCS" || op == "DECS" || op == "INCSCc" || op == "DECSCc") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i(), x = tk.i();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "INCC") {
                a->INCC(bitCapInt(v), s, l, x);
            } else if (op == "DECC") {
                a->DECC(bitCapInt(v), s, l, x);
            } else if (op == "INCS") {
                a->INCS(bitCapInt(v), s, l, x);
            } else if (op == "DECS") {
                a->DECS(bitCapInt(v), s, l, x);
            } else if (op == "INCSCc") {
                a->INCSC(bitCapInt(v), s, l, x);
            } else {
                a->DECSC(bitCapInt(v), s, l, x);
            }
        } else if (op == "INCSC" || op == "DECSC") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i(), o = tk.i(), cy = tk.i();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "INCSC") {
                a->INCSC(bitCapInt(v), s, l, o, cy);
            } else {
                a->DECSC(bitCapInt(v), s, l, o, cy);
            }
        } else if (op == "MUL" || op == "DIV") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), cs = tk.i(), l = tk.i();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "MUL") {
                a->MUL(bitCapInt(v), s, cs, l);
            } else {
                a->DIV(bitCapInt(v), s, cs, l);
            }
        } else if (op == "CMUL" || op == "CDIV") {
            std::vector<bitLenInt> c = tk.qubits();
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), cs = tk.i(), l = tk.i();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "CMUL") {
                a->CMUL(bitCapInt(v), s, cs, l, c);
            } else {
                a->CDIV(bitCapInt(v), s, cs, l, c);
            }
        } else if (op == "MULModNOut" || op == "IMULModNOut" || op == "POWModNOut") {
            const uint64_t v = tk.u(), n = tk.u();
            const bitLenInt is = tk.i(), os = tk.i(), l = tk.i();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "MULModNOut") {
                a->MULModNOut(bitCapInt(v), bitCapInt(n), is, os, l);
            } else if (op == "IMULModNOut") {
                a->IMULModNOut(bitCapInt(v), bitCapInt(n), is, os, l);
            } else {
                a->POWModNOut(bitCapInt(v), bitCapInt(n), is, os, l);
            }
        } else if (op == "CMULModNOut" || op == "CIMULModNOut" || op == "CPOWModNOut") {
            std::vector<bitLenInt> c = tk.qubits();
            const uint64_t v = tk.u(), n = tk.u();
            const bitLenInt is = tk.i(), os = tk.i(), l = tk.i();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "CMULModNOut") {
                a->CMULModNOut(bitCapInt(v), bitCapInt(n), is, os, l, c);
            } else if (op == "CIMULModNOut") {
                a->CIMULModNOut(bitCapInt(v), bitCapInt(n), is, os, l, c);
            } else {
                a->CPOWModNOut(bitCapInt(v), bitCapInt(n), is, os, l, c);
            }
        } else if (op == "IndexedLDA") {
            const bitLenInt is = tk.i(), il = tk.i(), vs = tk.i(), vl = tk.i();
            std::vector<unsigned char> tab = tk.hex();
            dynamic_cast<QAlu*>(q.get())->IndexedLDA(is, il, vs, vl, tab.data(), true);
        } else if (op == "IndexedADC" || op == "IndexedSBC") {
            const bitLenInt is = tk.i(), il = tk.i(), vs = tk.i(), vl = tk.i(), cy = tk.i();
            std::vector<unsigned char> tab = tk.hex();
            QAlu* a = dynamic_cast<QAlu*>(q.get());
            if (op == "IndexedADC") {
                a->IndexedADC(is, il, vs, vl, cy, tab.data());
            } else {
                a->IndexedSBC(is, il, vs, vl, cy, tab.data());
            }
        } else if (op == "Hash") {
            const bitLenInt s = tk.i(), l = tk.i();
            std::vector<unsigned char> tab = tk.hex();
            dynamic_cast<QAlu*>(q.get())->Hash(s, l, tab.data());
        } else if (op == "PhaseFlipIfLess") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i();
            dynamic_cast<QAlu*>(q.get())->PhaseFlipIfLess(bitCapInt(v), s, l);
        } else if (op == "CPhaseFlipIfLess") {
            const uint64_t v = tk.u();
            const bitLenInt s = tk.i(), l = tk.i(), f = tk.i();
            dynamic_cast<QAlu*>(q.get())->CPhaseFlipIfLess(bitCapInt(v), s, l, f);
#endif
        } else if (op == "SetPermutation") {
            q->SetPermutation(bitCapInt(tk.u()), ONE_CMPLX);
        } else if (op == "ForceM") {
            const bitLenInt t = tk.i();
            const bool res = tk.i() != 0;
            q->ForceM(t, res, true, true);
        } else if (op == "ForceMReg") {
            const bitLenInt s = tk.i(), l = tk.i();
            const uint64_t res = tk.u();
            q->ForceMReg(s, l, bitCapInt(res), true, true);
        } else if (op == "NormalizeState") {
            q->NormalizeState();
        } else if (op == "UpdateRunningNorm") {
            q->UpdateRunningNorm();
        } else if (op == "Compose") {
            const int src = (int)tk.i();
            if (tk.p < tk.t.size()) {
                const bitLenInt start = tk.i();
                q->Compose(regs.at(src), start);
            } else {
                q->Compose(regs.at(src));
            }
        } else if (op == "Decompose") {
            const bitLenInt s = tk.i(), l = tk.i();
            const int dst = (int)tk.i();
            regs[dst] = q->Decompose(s, l);
        } else if (op == "Dispose") {
            const bitLenInt s = tk.i(), l = tk.i();
            if (tk.p < tk.t.size()) {
                q->Dispose(s, l, bitCapInt(tk.u()));
            } else {
                q->Dispose(s, l);
            }
        } else if (op == "Allocate") {
            const bitLenInt s = tk.i(), l = tk.i();
            q->Allocate(s, l);
        } else if (op == "Prob") {
            snprintf(buf, sizeof(buf), "Prob %.17g", (double)q->Prob((bitLenInt)tk.i()));
            results.push_back(buf);
        } else if (op == "ProbAll") {
            snprintf(buf, sizeof(buf), "ProbAll %.17g", (double)q->ProbAll(bitCapInt(tk.u())));
            results.push_back(buf);
        } else if (op == "ProbReg") {
            const bitLenInt s = tk.i(), l = tk.i();
            snprintf(buf, sizeof(buf), "ProbReg %.17g", (double)q->ProbReg(s, l, bitCapInt(tk.u())));
            results.push_back(buf);
        } else if (op == "ProbMask") {
            const uint64_t mask = tk.u(), perm = tk.u();
            snprintf(buf, sizeof(buf), "ProbMask %.17g", (double)q->ProbMask(bitCapInt(mask), bitCapInt(perm)));
            results.push_back(buf);
        } else if (op == "ProbParity") {
            QParity* qp = dynamic_cast<QParity*>(q.get());
            snprintf(buf, sizeof(buf), "ProbParity %.17g", qp ? (double)qp->ProbParity(bitCapInt(tk.u())) : -1.0);
            results.push_back(buf);
        } else if (op == "CProb") {
            const bitLenInt c = tk.i(), t = tk.i();
            snprintf(buf, sizeof(buf), "CProb %.17g", (double)q->CProb(c, t));
            results.push_back(buf);
        } else if (op == "ACProb") {
            const bitLenInt c = tk.i(), t = tk.i();
            snprintf(buf, sizeof(buf), "ACProb %.17g", (double)q->ACProb(c, t));
            results.push_back(buf);
        } else if (op == "GetAmplitude") {
            const complex a = q->GetAmplitude(bitCapInt(tk.u()));
            snprintf(buf, sizeof(buf), "GetAmplitude %.17g %.17g", (double)real(a), (double)imag(a));
            results.push_back(buf);
        } else if (op == "SumSqrDiff") {
            const int other = (int)tk.i();
            snprintf(buf, sizeof(buf), "SumSqrDiff %.17g", (double)q->SumSqrDiff(regs.at(other)));
            results.push_back(buf);
        } else if (op == "Norm") {
            q->UpdateRunningNorm();
            QEnginePtr e = std::dynamic_pointer_cast<QEngine>(q);
            snprintf(buf, sizeof(buf), "Norm %.17g", e ? (double)e->GetRunningNorm() : -1.0);
            results.push_back(buf);
        } else {
            fprintf(stderr, "unknown op '%s'\n", op.c_str());
            return 3;
        }
    }
    if (timing) {
        for (auto& kv : regs) {
            kv.second->Finish();
        }
        const double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        seconds += dt;
        totalGates += gateCount;
        segments.push_back(std::make_pair(gateCount, dt));
    }

    if (!resultsFile.empty()) {
        std::ofstream out(resultsFile);
        for (const std::string& s : results) {
            out << s << "\n";
        }
    }
    if (!dumpPrefix.empty()) {
        for (auto& kv : regs) {
            QInterfacePtr q = kv.second;
            const size_t len = (size_t)1U << q->GetQubitCount();
            std::vector<complex> st(len);
            q->GetQuantumState(st.data());
            const std::string fn = dumpPrefix + "." + std::to_string(kv.first) + ".bin";
            FILE* f = fopen(fn.c_str(), "wb");
            fwrite(st.data(), sizeof(complex), len, f);
            fclose(f);
        }
    }
    if (doTime) {
        std::string seg = "[";
        for (size_t i = 0; i < segments.size(); ++i) {
            char b[64];
            snprintf(b, sizeof(b), "%s[%zu, %.6f]", i ? ", " : "", segments[i].first, segments[i].second);
            seg += b;
        }
        seg += "]";
        printf("{\"ops\": %zu, \"seconds\": %.6f, \"threads\": %u, \"fp_bits\": %d, \"engine\": \"%s\", \"segments\": %s}\n", totalGates,
            seconds, regs.empty() ? 0U : regs.begin()->second->GetConcurrencyLevel(), (int)(8 * sizeof(real1)), g_engine.c_str(),
            seg.c_str());
    }
    return 0;
}
