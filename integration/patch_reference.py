"""Build step of the resident front-end (integration/Makefile): writes patched copies of two reference sources into
integration/_build/ — the reference tree is read, never modified, and nothing of it is stored in this repository.

    python integration/patch_reference.py /root/reference integration/_build

Inserted (each right after the reference's own argument checks, so error behaviour is unchanged):

  ttNetwork.cpp  TTNetwork::round(maxRanks, eps)         if (xb200_resident::round(*this, _maxRanks, _eps)) return;
                 TTNetwork::move_core(position, keepRank) if (xb200_resident::move_core(*this, _position, _keepRank)) return;
                 TTNetwork::soft_threshold(taus, .)       if (xb200_resident::soft_threshold(*this, _taus)) return;
                 TTNetwork::operator+=(other)             if (xb200_resident::add(*this, _other)) return *this;
  als.cpp        ALSVariant::solve(...)                   { double e; if (xb200_resident::als_solve(*this, _Ap, _x, _b, ..., e)) return e; }
"""
import os
import re
import sys


def insert_after(src, signature_regex, anchor, hook, what):
    m = re.search(signature_regex, src)
    if not m:
        sys.exit("patch_reference: signature of %s not found" % what)
    pos = src.find(anchor, m.end())
    nxt = re.search(r"\n\ttemplate<bool isOperator>\n|\n\t[a-zA-Z].*::.*\) (const )?\{\n", src[m.end():])
    if pos < 0 or (nxt and pos > m.end() + nxt.start()):
        sys.exit("patch_reference: anchor of %s not found inside the function" % what)
    pos += len(anchor)
    return src[:pos] + "\n\t\t" + hook + "\n" + src[pos:]


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    inc = '#include "%s"\n' % os.path.join(os.path.dirname(os.path.abspath(__file__)), "xb200_resident.h")

    tt = open(os.path.join(ref, "src", "xerus", "ttNetwork.cpp")).read()
    tt = insert_after(tt, r"void TTNetwork<isOperator>::round\(const std::vector<size_t>& _maxRanks, const double _eps\) \{",
                      'REQUIRE(!misc::contains(_maxRanks, size_t(0)), "Trying to round a TTTensor to rank 0 is not possible.");',
                      "if (xb200_resident::round(*this, _maxRanks, _eps)) { return; }", "round")
    tt = insert_after(tt, r"void TTNetwork<isOperator>::move_core\(const size_t _position, const bool _keepRank\) \{",
                      "require_correct_format();", "if (xb200_resident::move_core(*this, _position, _keepRank)) { return; }", "move_core")
    tt = insert_after(tt, r"void TTNetwork<isOperator>::soft_threshold\(const std::vector<double> &_taus, const bool  /\*_preventZero\*/\) \{",
                      "require_correct_format();", "if (xb200_resident::soft_threshold(*this, _taus)) { return; }", "soft_threshold")
    tt = insert_after(tt, r"TTNetwork<isOperator>& TTNetwork<isOperator>::operator\+=\(const TTNetwork<isOperator>& _other\) \{",
                      "require_correct_format();", "if (xb200_resident::add(*this, _other)) { return *this; }", "operator+=")
    first = tt.index("#include")
    tt = tt[:first] + inc + tt[first:]
    open(os.path.join(out, "ttNetwork_resident.cpp"), "w").write(tt)

    als = open(os.path.join(ref, "src", "xerus", "algorithms", "als.cpp")).read()
    m = re.search(r"double ALSVariant::solve\(const TTOperator \*_Ap, TTTensor &_x, const TTTensor &_b, size_t _numHalfSweeps, value_t _convergenceEpsilon, PerformanceData &_perfData\) const \{", als)
    if not m:
        sys.exit("patch_reference: ALSVariant::solve not found")
    pos = als.find("#endif", m.end())
    if pos < 0:
        sys.exit("patch_reference: end of the argument checks of ALSVariant::solve not found")
    pos += len("#endif")
    hook = ("\n\t\t{ double xb200_energy = 0.0; if (xb200_resident::als_solve(*this, _Ap, _x, _b, _numHalfSweeps, _convergenceEpsilon, xb200_energy)) "
            "{ return xb200_energy; } }\n")
    als = als[:pos] + hook + als[pos:]
    first = als.index("#include")
    als = als[:first] + inc + als[first:]
    open(os.path.join(out, "als_resident.cpp"), "w").write(als)
    print("patched: ttNetwork_resident.cpp (round, move_core, soft_threshold, operator+=), als_resident.cpp (ALSVariant::solve)")


if __name__ == "__main__":
    main()
