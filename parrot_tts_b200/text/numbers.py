"""English number -> words, dependency-free.

The reference's cleaners call ``inflect.engine().number_to_words`` (reference utils/aligner/cleaners.py:5,93,100-117;
``inflect`` is an unpinned third-party dependency, requirements.txt, and is not installed here).  This module restates
the three call shapes the cleaners use, from inflect's published behaviour:

  * ``number_to_words(n, andword="")``                      cardinal, American style: groups of three digits joined by
    ", "; inside a group no "and"; when the LAST group is a single word ("one", "twenty-one") the comma before it is
    replaced by the and-word - i.e. by nothing: 124001 -> "one hundred twenty-four thousand one",
    1234 -> "one thousand, two hundred thirty-four";
  * ``number_to_words(n, andword="", zero="oh", group=2)``  digit pairs (years): 1901 -> "nineteen, oh one";
  * ``number_to_words("243rd")``                            ordinal with the default and-word:
    "two hundred and forty-third".

PARITY UNPINNED against inflect itself (absent); pinned by known answers in tests/test_text_cpu.py (the cases the
cleaners' lineage documents: years, ordinals, dollars, comma groups).
"""

_UNITS = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
          "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]
_MILL = ["", "thousand", "million", "billion", "trillion", "quadrillion", "quintillion", "sextillion", "septillion",
         "octillion", "nonillion", "decillion"]
_ORDINAL_WORD = {"one": "first", "two": "second", "three": "third", "five": "fifth", "eight": "eighth", "nine": "ninth",
                 "twelve": "twelfth"}


def _below_100(n):
    if n < 20:
        return _UNITS[n]
    t, u = divmod(n, 10)
    return _TENS[t] + ("-" + _UNITS[u] if u else "")


def _below_1000(n, andword):
    h, r = divmod(n, 100)
    if not h:
        return _below_100(r)
    if not r:
        return _UNITS[h] + " hundred"
    return " ".join(w for w in (_UNITS[h], "hundred", andword, _below_100(r)) if w)


def cardinal(n, andword=""):
    """Non-negative int -> words (three-digit groups)."""
    n = int(n)
    if n < 0:
        raise ValueError("cardinal: negative numbers never reach the cleaners (the regex matches digits only)")
    if n == 0:
        return "zero"
    groups = []
    while n:
        n, g = divmod(n, 1000)
        groups.append(g)
    if len(groups) > len(_MILL):
        raise ValueError("cardinal: number out of range (>= 10**36)")
    parts = []
    for i in range(len(groups) - 1, -1, -1):
        if groups[i]:
            parts.append(_below_1000(groups[i], andword) + (" " + _MILL[i] if i else ""))
    # the last three-digit group, when it is one word, is attached with the and-word instead of a comma
    if len(parts) > 1 and groups[0] and " " not in parts[-1]:
        tail = parts.pop()
        head = ", ".join(parts)
        return " ".join(w for w in (head, andword, tail) if w)
    return ", ".join(parts)


def digit_pairs(n, zero="oh"):
    """Digits read two at a time, left to right (inflect ``group=2``): 1901 -> 'nineteen, oh one'."""
    s = str(int(n))
    out = []
    for i in range(0, len(s), 2):
        pair = s[i:i + 2]
        if len(pair) == 1:
            out.append(_UNITS[int(pair)] if pair != "0" else zero)
        elif pair[0] != "0":
            out.append(_below_100(int(pair)))
        elif pair[1] != "0":
            out.append(zero + " " + _UNITS[int(pair[1])])
        else:
            out.append(zero + " " + zero)
    return ", ".join(out)


def ordinal(n, andword="and"):
    """int -> ordinal words: the last word of the cardinal takes the ordinal form."""
    words = cardinal(n, andword)
    cut = max(words.rfind(" "), words.rfind("-")) + 1
    head, last = words[:cut], words[cut:]
    if last in _ORDINAL_WORD:
        last = _ORDINAL_WORD[last]
    elif last.endswith("y"):
        last = last[:-1] + "ieth"
    else:
        last += "th"
    return head + last
