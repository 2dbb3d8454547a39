"""Text cleaners of the demo path, dependency-free.

Mirrors the three pipelines demo.ipynb cell 9 chooses between (reference utils/aligner/cleaners.py):

  english_cleaners                        :120-140   ASCII fold, lower case, numbers, abbreviations, blanks, character filter
  nonenglish_cleaners                     :142-167   ASCII transliteration, lower case, digits and symbols removed
  nonenglish_cleaners_no_transliteration  :205-233   symbols removed, Devanagari numerals spelt out in Hindi

The reference needs ``inflect`` (number words) and ``unidecode`` (transliteration); neither is installed here.  Number
words come from ``text.numbers``.  ``ascii_fold`` uses ``unidecode`` when it is importable and otherwise folds
Latin-script text only (compatibility decomposition, combining marks dropped, a table for the letters and punctuation
that do not decompose); ``nonenglish_cleaners`` - whose whole point is transliterating other scripts - raises
``ImportError`` without it instead of silently dropping the text.  Host string work: nothing here touches the device.
"""
import re
import unicodedata

from . import numbers as _num

# ---- ASCII folding ----------------------------------------------------------------------------------------
_FOLD = {
    "ß": "ss", "æ": "ae", "Æ": "AE", "œ": "oe", "Œ": "OE", "ø": "o", "Ø": "O", "đ": "d", "Đ": "D", "ð": "d", "Ð": "D",
    "þ": "th", "Þ": "Th", "ł": "l", "Ł": "L", "ı": "i", "ħ": "h", "Ħ": "H",
    "‘": "'", "’": "'", "‚": ",", "“": '"', "”": '"', "„": ",,", "–": "-", "—": "--",
    "…": "...", "«": "<<", "»": ">>", "‹": "<", "›": ">", " ": " ", "•": "*",
    "×": "x", "÷": "/", "€": "EUR", "£": "PS", "©": "(c)", "®": "(r)", "™": "(tm)",
    "°": "deg", "½": " 1/2", "¼": " 1/4", "¾": " 3/4",
}


def _have_unidecode():
    try:
        import unidecode  # noqa: F401
        return True
    except ImportError:
        return False


def ascii_fold(text, strict=False):
    """``unidecode.unidecode`` when available.  Otherwise: Latin-script folding; characters of other scripts are
    dropped (``strict=False``, which is what the English pipeline's printable-ASCII filter would do to whatever a
    transliteration left over anyway) or rejected (``strict=True``)."""
    if _have_unidecode():
        from unidecode import unidecode
        return unidecode(text)
    out = []
    for ch in text:
        if ord(ch) < 128:
            out.append(ch)
        elif ch in _FOLD:
            out.append(_FOLD[ch])
        else:
            base = "".join(c for c in unicodedata.normalize("NFKD", ch) if not unicodedata.combining(c))
            if base and all(ord(c) < 128 for c in base):
                out.append(base)
            elif all(c in _FOLD for c in base) and base:
                out.append("".join(_FOLD[c] for c in base))
            elif strict:
                raise ImportError(f"ascii_fold: {ch!r} (U+{ord(ch):04X}) needs the 'unidecode' package, which is not installed")
    return "".join(out)


# ---- numbers ----------------------------------------------------------------------------------------------
_RE_COMMA_NUMBER = re.compile(r"([0-9][0-9\,]+[0-9])")
_RE_DECIMAL = re.compile(r"([0-9]+\.[0-9]+)")
_RE_POUNDS = re.compile(r"£([0-9\,]*[0-9]+)")
_RE_DOLLARS = re.compile(r"\$([0-9\.\,]*[0-9]+)")
_RE_ORDINAL = re.compile(r"[0-9]+(st|nd|rd|th)")
_RE_NUMBER = re.compile(r"[0-9]+")
_RE_BLANKS = re.compile(r"\s+")


def _money(m):
    amount = m.group(1)
    pieces = amount.split(".")
    if len(pieces) > 2:
        return amount + " dollars"
    whole = int(pieces[0]) if pieces[0] else 0
    frac = int(pieces[1]) if len(pieces) == 2 and pieces[1] else 0
    said = []
    if whole:
        said.append(f"{whole} {'dollar' if whole == 1 else 'dollars'}")
    if frac:
        said.append(f"{frac} {'cent' if frac == 1 else 'cents'}")
    return ", ".join(said) if said else "zero dollars"


def _spoken_number(m):
    n = int(m.group(0))
    if 1000 < n < 3000:                       # read like a year
        if n == 2000:
            return "two thousand"
        if 2000 < n < 2010:
            return "two thousand " + _num.cardinal(n % 100)
        if n % 100 == 0:
            return _num.cardinal(n // 100) + " hundred"
        return _num.digit_pairs(n, zero="oh").replace(", ", " ")
    return _num.cardinal(n, andword="")


def _spoken_ordinal(m):
    return _num.ordinal(int(re.sub(r"(st|nd|rd|th)$", "", m.group(0))), andword="and")


def normalize_numbers(text):
    """cleaners.py:52-59: thousands separators, pounds, dollars, decimal points, ordinals, then plain numbers."""
    text = _RE_COMMA_NUMBER.sub(lambda m: m.group(1).replace(",", ""), text)
    text = _RE_POUNDS.sub(r"\1 pounds", text)
    text = _RE_DOLLARS.sub(_money, text)
    text = _RE_DECIMAL.sub(lambda m: m.group(1).replace(".", " point "), text)
    text = _RE_ORDINAL.sub(_spoken_ordinal, text)
    return _RE_NUMBER.sub(_spoken_number, text)


expand_numbers = normalize_numbers

# ---- abbreviations ----------------------------------------------------------------------------------------
_ABBREVIATIONS = [(re.compile(r"\b%s\." % a, re.IGNORECASE), full) for a, full in (
    ("mrs", "misess"), ("&", "and"), ("mr", "mister"), ("dr", "doctor"), ("st", "saint"), ("co", "company"),
    ("jr", "junior"), ("maj", "major"), ("gen", "general"), ("drs", "doctors"), ("rev", "reverend"),
    ("lt", "lieutenant"), ("hon", "honorable"), ("sgt", "sergeant"), ("capt", "captain"), ("esq", "esquire"),
    ("ltd", "limited"), ("col", "colonel"), ("ft", "fort"), ("tts", "text to speech"))]


def expand_abbreviations(text):
    for pat, full in _ABBREVIATIONS:
        text = pat.sub(full, text)
    return text


def collapse_whitespace(text):
    return _RE_BLANKS.sub(" ", text)


def lowercase(text):
    return text.lower()


def basic_cleaners(text):
    return collapse_whitespace(lowercase(text))


def _strip(text, chars):
    return text.translate({ord(c): None for c in chars})


_ENGLISH_DROP = "#+\\_`@/-'><()*\":;!"
_NONENGLISH_DROP = "|%+=[]^\\{}_`‘’@/-'><()*\":;!"
_NATIVE_DROP = "—⇒'–…“”%+=[]^\\{}_`‘’@/-><()*\":;!"
_NATIVE_CONTROL = "\x92\xad\x10\x13\x14\x16\x91"


def english_cleaners(text):
    text = ascii_fold(text).lower()
    text = collapse_whitespace(expand_abbreviations(normalize_numbers(text)))
    text = re.sub(r"[^\x20-\x7E]", "", text)
    return _strip(text, _ENGLISH_DROP).replace("&", "and")


def nonenglish_cleaners(text):
    if not _have_unidecode() and any(ord(c) > 127 and unicodedata.category(c)[0] == "L"
                                     and not unicodedata.name(c, "").startswith("LATIN") for c in text):
        raise ImportError("nonenglish_cleaners transliterates other scripts with 'unidecode', which is not installed; "
                          "use nonenglish_cleaners_no_transliteration (demo.ipynb: transliteration = False)")
    text = collapse_whitespace(ascii_fold(text).lower())
    text = _strip(text, "0123456789")
    text = re.sub(r"[^\x20-\x7E]", "", text)
    text = _strip(text, _NONENGLISH_DROP).replace("&", "and")
    return " ".join(text.split())


# ---- Hindi numerals ---------------------------------------------------------------------------------------
hindi_numbers = {
    0: "शून्य", 1: "एक", 2: "दो", 3: "तीन", 4: "चार", 5: "पाँच", 6: "छह", 7: "सात", 8: "आठ", 9: "नौ",
    10: "दस", 11: "ग्यारह", 12: "बारह", 13: "तेरह", 14: "चौदह", 15: "पंद्रह", 16: "सोलह", 17: "सत्रह", 18: "अठारह",
    19: "उन्नीस", 20: "बीस", 30: "तीस", 40: "चालीस", 50: "पचास", 60: "साठ", 70: "सत्तर", 80: "अस्सी", 90: "नब्बे",
}
hindi_numbers.update({100 * k: (hindi_numbers[k] + " " if k > 1 else "") + "सौ" for k in range(1, 10)})


def number_to_hindi(num):
    """cleaners.py:169-185: table look-up, tens + ones, hundreds + rest; ``None`` from 1000 up, like the reference."""
    if num in hindi_numbers:
        return hindi_numbers[num]
    if num < 100:
        return hindi_numbers[num - num % 10] + " " + hindi_numbers[num % 10]
    if num < 1000:
        return hindi_numbers[num - num % 100] + " " + number_to_hindi(num % 100)
    return None


_RE_DEVANAGARI_DIGITS = re.compile("[०-९]+")


def replace_devanagari_numbers(text):
    def spell(m):
        return number_to_hindi(int("".join(str(ord(c) - ord("०")) for c in m.group(0))))
    return _RE_DEVANAGARI_DIGITS.sub(spell, text)


def nonenglish_cleaners_no_transliteration(text):
    text = collapse_whitespace(text)
    text = _strip(text, "0123456789")
    text = _strip(text, _NATIVE_DROP).replace("|", ".")
    text = replace_devanagari_numbers(text)
    text = _strip(text, _NATIVE_CONTROL)
    return " ".join(text.split())
