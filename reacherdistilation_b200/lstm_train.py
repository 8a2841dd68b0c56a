"""LSTM distillation entry point (/root/reference src/distilation/lstm_train.py:18-201).

SURVEY 8(f) ranks the LSTM student as the first row AFTER the MLP hot path meets its bar; it is not built yet.  The entry
point exists so `main.py -lt` fails loudly instead of silently doing something else."""


def train(train, restore):
    raise NotImplementedError("LSTM student (student_nn.py:21-49) is scheduled after the MLP hot path -- see DESIGN.md 'next'")
