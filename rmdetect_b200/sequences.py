"""Input sequences as the scan's callers hand them over: the reference's `Sequences` / `Sequence` objects and its readers.

Restates, for a caller that has no reference checkout, what `rmdetect.py:228-236` builds before it
calls `Scanner.scan` (paths relative to the reference root):

  lib/sequences.py:7-64     Sequences: the loader is chosen by the file name's ending (.seq, .fasta, .stk / .stockholm,
                            .aln), a string or stdin gives one sequence called "unique"; `both_strands` appends the
                            reverse complements (key + "--") after all forward sequences (:103-111)
  lib/sequences.py:114-167  Sequence: strip, drop newlines, upper case, T -> U (:145-150); a sequence with one of the
                            gap characters ".-_" keeps its gapped form and a column index (:157-167), which
                            `Scanner.slide` uses for `chains_cols` (lib/scanner.py:100-111)
  lib/file_io.py:5-24       Fasta: key = first word of the header line, sequence lines concatenated
  lib/file_io.py:209-316    Stk (Stockholm): blocks of one key are concatenated, rows with a character outside
                            "ACUG" and the gap characters are skipped, the gap characters "-_:,~" become ".",
                            an unbalanced `#=GC SS_cons` line rejects the whole file (no sequences)
  lib/file_io.py:553-577    Clustal: rows after the CLUSTAL line, conservation lines (those with a "*") skipped

Python-3 notes: the reference's Clustal.add_seq calls `dict.has_key` (lib/file_io.py:480) and fails under Python 3; the
reader here does what that line means.  Nothing here touches the GPU: the hot path starts at `Scanner.scan`.
"""
import os
import sys

VALID_GAPS = ".-_"                                    # lib/sequences.py:115

STK_NUCLEOTIDES = "ACUG"                              # lib/file_io.py:40-44
STK_GAP_STANDARD = "."
STK_GAP_NON_STANDARD = "-_:,~"
STK_CHAR_ALLOWED = STK_NUCLEOTIDES + STK_GAP_STANDARD + STK_GAP_NON_STANDARD
STK_OPEN = "(<{[ABCDEFGHIJ"                           # lib/file_io.py:46-47
STK_CLOSE = ")>}]abcdefghij"


class Sequence:
    """One input sequence (lib/sequences.py:114-167): `seq` without gaps, `seq_gapped` as read, `col_index[k]` = alignment
    column of letter k (None when the input had no gap character)."""

    VALID_GAPS = VALID_GAPS

    def __init__(self, key="", seq="", rev=False):
        self.key = key
        self.seq = seq.strip().replace("\n", "").upper().replace("T", "U")      # :145-150
        self.seq_gapped = self.seq
        self.rev = rev
        self.col_index = None
        if any(gap in self.seq for gap in VALID_GAPS):                          # :127-135
            cols = [n for n, c in enumerate(self.seq_gapped) if c not in VALID_GAPS]
            self.seq = "".join(self.seq_gapped[n] for n in cols)
            self.col_index = cols
        self.seq_len = len(self.seq)

    def reverse(self):
        """The reverse complement as a sequence of its own, key + "--" (:139-143).  The reference complements the UNGAPPED
        letters and builds the new object from them, so a reverse strand never has a column index."""
        comp = {"A": "U", "C": "G", "G": "C", "U": "A"}
        return Sequence(self.key + "--", "".join(comp.get(c, c) for c in self.seq[::-1]), True)


def read_fasta(fi):
    """[(key, sequence)] of a FASTA stream (lib/file_io.py:12-24): blank lines ignored, key = first word after '>'."""
    seqs = []
    for line in fi:
        line = line.strip()
        if line == "":
            continue
        if line.startswith(">"):
            seqs.append([line[1:].split()[0], ""])
        else:
            seqs[-1][1] += line                     # a sequence line before any header fails here, as in the reference
    return [(k, s) for k, s in seqs]


def _add_row(seqs, index, key, row):
    """A key's rows are concatenated in file order (lib/file_io.py:73-91, :478-496)."""
    if key not in index:
        index[key] = len(seqs)
        seqs.append((key, row))
    else:
        i = index[key]
        seqs[i] = (key, seqs[i][1] + row)


def read_stockholm(fi, char_allowed_extra="", warn=sys.stderr):
    """[(key, gapped sequence)] of a Stockholm stream (lib/file_io.py:209-316); [] when the file is rejected."""
    seqs, index = [], {}
    secondary = ""
    header_ok = end_ok = ss_ok = False
    for count, line in enumerate(fi, 1):
        line = line.strip()
        if line == "":
            continue
        if line.startswith("# STOCKHOLM 1.0"):
            header_ok = True
            continue
        if not header_ok:                            # everything before the signature is ignored
            continue
        if line == "//":
            end_ok = True
            break
        data = line.split()
        if line.startswith("#=GC SS_cons "):
            secondary += data[2]
            ss_ok = True
        elif line.startswith("#"):                   # #=GC RF, other #=GC rows, #=GF / #=GS / #=GR: not sequences
            continue
        else:
            if len(data) > 2:
                if warn:
                    warn.write("Error in line %d -> No spaces allowed in key or sequence.\n" % count)
                return []
            row = data[1].upper()
            bad = [c for c in row if c not in STK_CHAR_ALLOWED + char_allowed_extra]
            if bad:
                if warn:
                    warn.write("Warning (line %d) -> Invalid character(s): '%s'. Skipping line.\n" % (count, ",".join(bad)))
                continue
            for c in STK_GAP_NON_STANDARD:
                row = row.replace(c, STK_GAP_STANDARD)
            _add_row(seqs, index, data[0], row)
    # an unbalanced consensus structure rejects the file (:282-304)
    stacks = [0] * len(STK_OPEN)
    for pos, s in enumerate(secondary, 1):
        i = STK_OPEN.find(s)
        if i > -1:
            stacks[i] += 1
            continue
        i = STK_CLOSE.find(s)
        if i > -1:
            if stacks[i] == 0:
                if warn:
                    warn.write("Error (sec. struct. pos %d) -> Unmatched closing char: %s\n" % (pos, s))
                return []
            stacks[i] -= 1
    for i, open_count in enumerate(stacks):
        if open_count > 0:
            if warn:
                warn.write("Error (sec. struct.) -> Not enough closing chars for '%s'.\n" % STK_OPEN[i])
            return []
    if warn:
        if not header_ok:
            warn.write("Warning -> '# STOCKHOLM 1.0' signature not found.\n")
        if not end_ok:
            warn.write("Warning -> End line string, '//', not found.\n")
        if not ss_ok:
            warn.write("Warning -> Secondary structure not found.\n")
    return seqs


def read_clustal(fi, warn=sys.stderr):
    """[(key, gapped sequence)] of a Clustal stream (lib/file_io.py:553-577)."""
    seqs, index = [], {}
    clustal_ok = False
    for line in fi:
        line = line.strip()
        if line == "":
            continue
        if not clustal_ok and line.startswith("CLUSTAL"):
            clustal_ok = True
        elif clustal_ok and line.find("*") >= 0:     # conservation line
            continue
        elif clustal_ok:
            data = line.split()
            _add_row(seqs, index, data[0], data[1])
        else:
            if warn:
                warn.write("ERROR: Not a clustal file\n")
            break
    return seqs


class Sequences:
    """All sequences of one input (lib/sequences.py:7-111), in the order `Scanner.scan` walks them."""

    def __init__(self, fname=None, seq=None, both_strands=False, stdin=None):
        self.fname = fname
        self.from_stdin = False
        self.alignment = False
        self.both_strands = both_strands
        self.sequence_list = []
        self.sequence_max_pos = 0
        self.sequence_max_col = 0
        self.sequence_index = {}
        if seq is not None:
            self._add("unique", seq)
            self.from_stdin = True
        elif fname is not None:
            if not os.path.isfile(fname):
                sys.stderr.write("File not found: %s\n" % fname)
                raise SystemExit                     # the reference calls quit()
            with open(fname) as fi:
                if fname.endswith(".seq"):
                    self._add("unique", fi.read())
                elif fname.endswith(".fasta"):
                    self._add_all(read_fasta(fi))
                elif fname.endswith(".stk") or fname.endswith(".stockholm"):
                    self._add_all(read_stockholm(fi))
                    self.alignment = True
                elif fname.endswith(".aln"):
                    self._add_all(read_clustal(fi))
                    self.alignment = True
                # any other ending: no sequences, as in the reference
        else:
            self._add("unique", (stdin or sys.stdin).read())
            self.from_stdin = True
        for s in self.sequence_list:                 # :57-59
            self.sequence_max_pos = max(self.sequence_max_pos, len(s.seq))
            self.sequence_max_col = max(self.sequence_max_col, len(s.seq_gapped), len(s.seq))
        if self.both_strands:                        # :103-111: all reverse strands after all forward ones
            for r in [s.reverse() for s in self.sequence_list]:
                self.sequence_index[r.key] = len(self.sequence_list)
                self.sequence_list.append(r)

    def _add(self, key, seq):
        self.sequence_index[key] = len(self.sequence_list)
        self.sequence_list.append(Sequence(key, seq))

    def _add_all(self, pairs):
        for key, seq in pairs:
            self._add(key, seq)

    def seq_count(self):
        """The reference's count (:66-70): it doubles an already doubled list when both strands are on."""
        return len(self.sequence_list) * 2 if self.both_strands else len(self.sequence_list)

    def get_sequence(self, key):
        return self.sequence_list[self.sequence_index[key]]
