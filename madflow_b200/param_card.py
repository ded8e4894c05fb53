"""Minimal SLHA param_card reader with the slice of MG5's `check_param_card.ParamCard` interface
that the generated `get_model_param` uses: card['BLOCK'].get(code).value
(reference: template_files/matrix_method_python.inc:30,37; PyOut_exporter.py:198-201)."""


class _Entry:
    def __init__(self, value):
        self.value = value


class _Block(dict):
    def get(self, code, default=None):
        if isinstance(code, (tuple, list)):
            code = code[0] if len(code) == 1 else tuple(code)
        return dict.get(self, code, default)


class ParamCard(dict):
    def __init__(self, path):
        super().__init__()
        block = None
        for raw in open(path):
            line = raw.split("#")[0].strip()
            if not line:
                continue
            low = line.lower()
            if low.startswith("block"):
                block = _Block()
                self[line.split()[1].upper()] = block
                self[line.split()[1].lower()] = block
                self[line.split()[1]] = block
            elif low.startswith("decay"):
                parts = line.split()
                for key in ("DECAY", "decay"):
                    self.setdefault(key, _Block())[int(parts[1])] = _Entry(float(parts[2]))
                block = None
            elif block is not None:
                parts = line.split()
                try:
                    codes = tuple(int(p) for p in parts[:-1])
                    block[codes[0] if len(codes) == 1 else codes] = _Entry(float(parts[-1]))
                except ValueError:
                    continue
