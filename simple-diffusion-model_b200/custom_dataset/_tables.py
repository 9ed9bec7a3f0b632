"""Reader for the TinyDB files the reference's labelled datasets use (custom_dataset/conditional_img_dataset.py:18-33).
A TinyDB database is a JSON document {table: {doc_id: row}}, so `json` is enough -- no tinydb dependency."""
import json


def load_tables(path):
    with open(path, "r") as f:
        db = json.load(f)
    data = list(db.get("Data", {}).values())
    labels = list(db.get("Labels", {}).values())
    if not data:
        raise Exception("No data found in Data table.")
    if not labels:
        raise Exception("No data found in Labels table.")
    return data, labels[0]["labels"]
