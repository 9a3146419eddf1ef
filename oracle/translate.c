/*
 * translate.c -- ORACLE (test infrastructure): six-frame translation and reverse
 * complement with Biopython semantics, as used by the reference at
 * /root/reference/lib/aln.py:241-258 (Bio.Seq.translate / reverse_complement;
 * Biopython is not vendored).
 */
#include "p7oracle.h"
#include <ctype.h>
#include <string.h>

/* NCBI genetic codes, codon order TTT TTC TTA TTG TCT ... GGG (base order T C A G) */
static const struct { int id; const char *aa; } TABLES[] = {
    {1, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {2, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG"},
    {3, "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {4, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {5, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG"},
    {6, "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {9, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
    {10, "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {11, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {12, "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {13, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG"},
    {14, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
    {15, "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {16, "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {21, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
    {22, "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {23, "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {24, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG"},
    {25, "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {26, "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {0, NULL}};

int orc_codon_table(int table, char *out64) {
  for (int i = 0; TABLES[i].aa; i++)
    if (TABLES[i].id == table) {
      memcpy(out64, TABLES[i].aa, 64);
      return 0;
    }
  return -1;
}

/* IUPAC nucleotide -> bitmask over T(1) C(2) A(4) G(8); 0 = invalid */
static int iupac_mask(char c) {
  switch (toupper((unsigned char)c)) {
    case 'T': case 'U': return 1;
    case 'C': return 2;
    case 'A': return 4;
    case 'G': return 8;
    case 'R': return 4 | 8;
    case 'Y': return 1 | 2;
    case 'M': return 4 | 2;
    case 'K': return 8 | 1;
    case 'S': return 2 | 8;
    case 'W': return 4 | 1;
    case 'H': return 4 | 2 | 1;
    case 'B': return 2 | 8 | 1;
    case 'V': return 4 | 2 | 8;
    case 'D': return 4 | 8 | 1;
    case 'N': return 15;
    default: return 0;
  }
}

int orc_translate(const char *nt, int n, int table, char *out) {
  char aa[64];
  if (orc_codon_table(table, aa) != 0) return -2;
  int naa = 0;
  for (int i = 0; i + 3 <= n; i += 3) {
    int m0 = iupac_mask(nt[i]), m1 = iupac_mask(nt[i + 1]), m2 = iupac_mask(nt[i + 2]);
    if (!m0 || !m1 || !m2) return -1; /* Biopython raises TranslationError */
    int nstop = 0, ntot = 0;
    unsigned seen = 0; /* bitset over 'A'..'Z' */
    for (int a = 0; a < 4; a++) if (m0 & (1 << a))
      for (int b = 0; b < 4; b++) if (m1 & (1 << b))
        for (int c = 0; c < 4; c++) if (m2 & (1 << c)) {
          char r = aa[a * 16 + b * 4 + c];
          ntot++;
          if (r == '*') nstop++;
          else seen |= 1u << (r - 'A');
        }
    char r;
    if (nstop == ntot) r = '*';
    else if (nstop > 0) r = 'X';
    else if ((seen & (seen - 1)) == 0) { r = 'A'; while (!(seen & (1u << (r - 'A')))) r++; }
    else if ((seen & ~((1u << ('D' - 'A')) | (1u << ('N' - 'A')))) == 0) r = 'B';
    else if ((seen & ~((1u << ('E' - 'A')) | (1u << ('Q' - 'A')))) == 0) r = 'Z';
    else if ((seen & ~((1u << ('I' - 'A')) | (1u << ('L' - 'A')))) == 0) r = 'J';
    else r = 'X';
    out[naa++] = r;
  }
  return naa;
}

void orc_revcomp(const char *nt, int n, char *out) {
  static const char *from = "ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn";
  static const char *to = "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn";
  for (int i = 0; i < n; i++) {
    char c = nt[n - 1 - i];
    const char *q = c ? strchr(from, c) : NULL;
    out[i] = q ? to[q - from] : c;
  }
}
